"""DRAM traffic of ONE whole training step measured over a profiler range, so that the two overlapped backward launches are
NOT serialised (ncu --replay-mode app-range re-runs the program once per metric pass):

  ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none --csv --log-file out.csv python scripts/range_traffic.py [stream_sms]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robust_nerf_b200 as rn                                   # noqa: E402
from robust_nerf_b200 import _lib                               # noqa: E402

lib = _lib.lib()
if len(sys.argv) > 1:
    lib.rn_set_flag(9, int(sys.argv[1]))
if len(sys.argv) > 2:
    lib.rn_set_flag(11, int(sys.argv[2]))
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    # set aside part of L2 for lines written / read with the evict_last (persisting) priority
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    mx = ctypes.c_int(0)
    rt.cudaDeviceGetAttribute(ctypes.byref(mx), ctypes.c_int(108), ctypes.c_int(0))      # cudaDevAttrMaxPersistingL2CacheSize
    want = min(int(sys.argv[3]) << 20, mx.value)
    rc = rt.cudaDeviceSetLimit(ctypes.c_int(0x06), ctypes.c_size_t(want))
    got = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(got), ctypes.c_int(0x06))
    rt.cudaGetLastError()
    print("cudaLimitPersistingL2CacheSize: max", mx.value >> 20, "MB, rc", rc, "now", got.value >> 20, "MB", file=sys.stderr)
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
for _ in range(3):
    trainer.step_rays(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer.step_rays(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
