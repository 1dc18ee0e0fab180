"""Stress the chain -> stream hand-off: the same backward pass N times with the weight gradients beside the data-gradient
chain; every run must be bit-identical to the first (a block read before its data had landed, or a lost flag, would show up
as a different sum), and within fp32 summation-order distance of the split-K result.  usage: python scripts/stress_stream.py [N]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import robust_nerf_b200 as rn                                   # noqa: E402
from robust_nerf_b200 import _lib                               # noqa: E402
from oracle import nerf_oracle as O                             # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
lib = _lib.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = rn.NeRF().to(dev)
sd = net.state_dict()
for k, v in O.make_weights(17, sharpen=True).items():
    sd[k] = torch.from_numpy(v).to(dev)
net.load_state_dict(sd)
rng = np.random.default_rng(3)
worst = 0.0
for M in (786432, 262144, 100003):
    pts = torch.as_tensor(rng.uniform(-3, 3, (M, 3)).astype(np.float32), device=dev)
    dirs = torch.as_tensor(rng.standard_normal((M, 3)).astype(np.float32), device=dev)
    gout = torch.as_tensor(rng.standard_normal((M, 4)).astype(np.float32), device=dev)

    def grads():
        net.zero_grad()
        raw = net.forward_raw(pts, dirs, 1)
        (raw * gout).sum().backward()
        return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()

    lib.rn_set_flag(9, 0)
    ref = grads()
    lib.rn_set_flag(9, 84)
    first = grads()
    rel = ((first - ref).abs().max() / ref.abs().max()).item()
    worst = max(worst, rel)
    bad = 0
    for i in range(N):
        g = grads()
        if not torch.equal(g, first):
            bad += 1
            print(f"M={M} run {i}: differs from the first stream run by {(g - first).abs().max().item():.3e}")
    torch.cuda.synchronize()
    print(f"M={M}: {N} stream runs, {bad} not bit-identical; stream vs split-K max diff / max entry = {rel:.2e}")
    assert bad == 0 and rel < 1e-3
print("stress ok, worst relative difference to the split-K result", worst)
