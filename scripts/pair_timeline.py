#!/usr/bin/env python
"""Timeline of the CTA-pair chain kernel (RN_EXPERIMENTS build): per-role event times of cluster 0 for a few items.
    RN_EXPERIMENTS=1 python robust-nerf_b200/build.py && python scripts/pair_timeline.py [train|bwd]
"bwd" records the data-gradient chain (mlp_chain_pair_bwd_kernel; role 3 is then the store warp)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn
from robust_nerf_b200 import _lib
lib = _lib.lib()
cdll = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
dev = torch.device("cuda:0")
torch.manual_seed(42)
coarse, fine = rn.create_nerf()
fine = fine.to(dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
train = mode in ("train", "bwd")
M = 148 * 2 * 256 * (4 if mode == 'bwd' else 8)           # 16 pair tiles per cluster
pts = torch.randn(M, 3, device=dev)
dirs = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
roles, entries = ctypes.c_int(), ctypes.c_int()
cdll.rn_pair_timeline_dims(ctypes.byref(roles), ctypes.byref(entries))
buf = torch.zeros(2 * roles.value * entries.value * 3, dtype=torch.int64, device=dev)
def run():
    if mode == "bwd":
        raw = fine.forward_raw(pts, dirs, 1)
        raw.backward(torch.ones_like(raw))
        return raw
    if train:
        x = pts.clone().requires_grad_(True)
        return fine.forward_raw(x, dirs, 1)
    with torch.no_grad():
        return fine.forward_raw(pts, dirs, 1)
run(); torch.cuda.synchronize()
setter = cdll.rn_pair_timeline_bwd if mode == "bwd" else cdll.rn_pair_timeline
setter.argtypes = [ctypes.c_void_p]
setter(ctypes.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
setter(ctypes.c_void_p(0))
t = buf.cpu().numpy().reshape(2, roles.value, entries.value, 3)
names = {0: "mma", 1: "epi_w2", 2: "epi_w17", 3: "store" if mode == "bwd" else "load"}
out = os.path.join(ROOT, "gpurun_out", "pair_timeline_%s.txt" % mode)
with open(out, "w") as fh:
    for rank in range(2):
        t0 = min(int(t[rank, r, 0, 1]) for r in range(roles.value) if t[rank, r, 0, 1] > 0)
        ev = []
        for r in range(roles.value):
            for e in range(entries.value):
                tag, clk, g = (int(v) for v in t[rank, r, e])
                if clk == 0: break
                ev.append((clk - t0, names[r], tag, g))
        ev.sort()
        for clk, nm, tag, g in ev[:1400]:
            fh.write(f"rank{rank} {clk:9d} {nm:8s} {tag:5d} {g}\n")
print("wrote", out)
