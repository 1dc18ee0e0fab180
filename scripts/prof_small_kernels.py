#!/usr/bin/env python
"""Launch the HBM-bound kernels once each at the config-5 stress size (65,536 rays, 128+256 samples) for ncu."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn
from robust_nerf_b200 import ops
from robust_nerf_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, Nc, Nf = 65536, 128, 256
ro = torch.randn(B, 3, device=dev, generator=g)
rd = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g), dim=-1)
zb = torch.linspace(2.0, 6.0, Nc, device=dev)
t_rand = torch.rand(B, Nc, device=dev, generator=g)
w = torch.rand(B, Nc, device=dev, generator=g) ** 4
u = torch.rand(B, Nf, device=dev, generator=g)
S = Nc + Nf
raw = torch.randn(B, S, 4, device=dev, generator=g); raw[..., 3] = raw[..., 3].abs() * 10
zz = torch.sort(torch.rand(B, S, device=dev, generator=g) * 4 + 2, -1)[0]
outs = [torch.empty(B, 3, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev), torch.empty(B, S, device=dev)]
d_raw = torch.empty(B, S, 4, device=dev)
gm = torch.randn(B, 3, device=dev, generator=g)
for _ in range(3):
    z, _p = ops.stratified(ro, rd, zb, t_rand)
    ops.sample_hierarchical(ro, rd, z, w, u)
    call("rn_composite_fwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, 0.0, ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), stream_ptr())
    call("rn_composite_bwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, ptr(gm), None, None, None, None, None, ptr(d_raw), None, stream_ptr())
torch.cuda.synchronize()
print("ok")
