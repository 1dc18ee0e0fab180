"""Hand-off lag of the weight-gradient stream (publication of a block by the data-gradient chain -> issue of its load),
for one backward pass of the fine network (786,432 points) and one of the coarse (262,144).  usage: python scripts/diag_lag.py"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn                                   # noqa: E402
from robust_nerf_b200 import _lib                               # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = rn.NeRF().to(dev)
rng = np.random.default_rng(3)
for stagger, sms, mode in ((0, 84, 32), (0, 84, 35)):      # 32: beside the chain; 35: after it (alone, every flag set)
    lib.rn_set_flag(11, stagger)
    lib.rn_set_flag(9, sms)
    for M in (786432,):
        pts = torch.as_tensor(rng.uniform(-3, 3, (M, 3)).astype(np.float32), device=dev)
        dirs = torch.as_tensor(rng.standard_normal((M, 3)).astype(np.float32), device=dev)
        gout = torch.as_tensor(rng.standard_normal((M, 4)).astype(np.float32), device=dev)
        for rep in range(3):
            lib.rn_set_flag(10, mode if rep == 2 else (mode & 3))
            net.zero_grad()
            raw = net.forward_raw(pts, dirs, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            (raw * gout).sum().backward()
            e1.record()
            torch.cuda.synchronize()
        mean, mx, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
        lib.rn_debug_stream_lag(ctypes.byref(mean), ctypes.byref(mx), ctypes.byref(n))
        busy = (ctypes.c_uint * 168)()
        lib.rn_debug_stream_busy(busy, 168)
        # pairs in problem order (dir, feature, L7 .. L1, L0), 4 or 5 pairs per problem at 84 SMs
        print("busy us per pair (leader CTAs):", [busy[2 * i] for i in range(42)])
        print(f"mode {mode}, stream SMs {sms}, M={M}: backward {e0.elapsed_time(e1):.3f} ms; hand-off lag mean {mean.value:.1f} us, max {mx.value:.1f} us over {n.value} CTAs")
lib.rn_set_flag(10, 0); lib.rn_set_flag(11, 0); lib.rn_set_flag(9, 84)
