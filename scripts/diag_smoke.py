#!/usr/bin/env python
"""smoke()-sized training step under every forward / backward kernel variant, against the fp32 and the bf16-emulating oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn
from robust_nerf_b200 import _lib
from oracle import nerf_oracle as O
lib = _lib.lib()
dev = torch.device("cuda:0")
wc, wf = O.make_weights(21), O.make_weights(22)
rng = np.random.default_rng(0)
B = 64
poses = rn.lego_poses(); H = W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
flat = rng.integers(0, H * W, B)
dirs = O.get_ray_directions(H, W, focal).reshape(-1, 3)[flat]
ro, rd = O.get_rays(dirs, poses[7].numpy())
target = rng.uniform(0, 1, (B, 3)).astype(np.float32)
t_rand = rng.uniform(0, 1, (B, 64)).astype(np.float32)
u = rng.uniform(0, 1, (B, 128)).astype(np.float32)
refs = {e: O.train_step_grads(wc, wf, ro, rd, target, t_rand=t_rand, u=u, emulate_bf16=e) for e in (False, True)}
for fwd, bwd in ((0, 0), (1, 0), (2, 0), (2, 1)):
    lib.rn_set_flag(0, fwd); lib.rn_set_flag(3, bwd)
    nets = []
    for w in (wc, wf):
        net = rn.NeRF().to(dev); sd = net.state_dict()
        for k, v in w.items(): sd[k] = torch.from_numpy(v).to(dev)
        net.load_state_dict(sd); nets.append(net)
    T = lambda a: torch.from_numpy(a).to(dev)
    out = rn.render_rays(nets[0], nets[1], T(ro), T(rd), rn.RenderConfig(), is_train=True, t_rand=T(t_rand), u=T(u))
    loss = ((out["rgb_coarse"] - T(target)) ** 2).mean() + ((out["rgb_fine"] - T(target)) ** 2).mean()
    loss.backward(); torch.cuda.synchronize()
    msg = f"fwd={fwd} bwd={bwd}:"
    for e in (False, True):
        ref = refs[e]
        err = float(np.abs(out["rgb_fine"].detach().cpu().numpy() - ref["rgb_fine"]).max())
        rels = []
        for name in ("pts_linears.0.weight", "pts_linears.3.weight", "pts_linears.7.weight", "dir_linear.weight"):
            mod = nets[1]
            for part in name.split("."): mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
            g = mod.grad.cpu().numpy(); gr = ref["grads_fine"][name]
            rels.append(float(np.linalg.norm(g - gr) / max(np.linalg.norm(gr), 1e-20)))
        msg += f"  [{'bf16-emu' if e else 'fp32'} oracle] rgb err {err:.2e} grad rel " + " ".join(f"{r:.3f}" for r in rels)
    print(msg, flush=True)
lib.rn_set_flag(0, 2); lib.rn_set_flag(3, 1)
