"""Gradient of ONE training step with the weight gradients after the chain (flag 9 = 0) against beside it (flag 9 = N):
per-parameter largest difference relative to the tensor's largest entry.  usage: python scripts/diag_stream.py [N]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robust_nerf_b200 as rn                                   # noqa: E402
from robust_nerf_b200 import _lib                               # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 72
lib = _lib.lib()
dev = torch.device("cuda", 0)
scene = rn.make_scene(800, 800, 100, seed=0, device=dev)
torch.manual_seed(42)
coarse, fine = rn.create_nerf(rn.ModelConfig())
coarse, fine = coarse.to(dev), fine.to(dev)
trainer = rn.Trainer(coarse, fine, rn.RenderConfig(), lr=5e-4)
ds, sampler = rn.create_pixel_dataset(scene)
g = torch.Generator(device="cpu").manual_seed(42)
idx = torch.randint(0, ds.n_pixels, (4096,), generator=g).to(dev)
pb = sampler.batch_from_indices(idx)
with torch.no_grad():
    ro, rd = sampler.get_rays_for_batch(pb, scene.poses)
batch = (ro.contiguous(), rd.contiguous(), pb.target_rgb.contiguous())


def grads(flag):
    lib.rn_set_flag(9, flag)
    torch.manual_seed(7)
    loss = trainer.step_rays(*batch, optimise=False)
    torch.cuda.synchronize()
    return float(loss), trainer.gflat.clone()


l0, g0 = grads(0)
l1, g1 = grads(N)
l2, g2 = grads(N)
l3, g3 = grads(0)
print("loss", l0, l1, l2, l3, "seq repeat equal", torch.equal(g0, g3), "stream repeat equal", torch.equal(g1, g2))
names = []
for tag, m in (("fine", fine), ("coarse", coarse)):
    for n, p in m.named_parameters():
        names.append((tag + "." + n, trainer._param_off[id(p)], p.numel()))
for name, off, n in names:
    a, b = g0[off:off + n], g1[off:off + n]
    d = (a - b).abs().max().item()
    print(f"{name:32s} max|a| {a.abs().max().item():.4e} max|diff| {d:.3e} rel {d / max(a.abs().max().item(), 1e-30):.2e}")
