#!/usr/bin/env python
"""Diagnose rays whose rgb differs by more than 1e-2 between the CUDA path and the fp32 PyTorch restatement on weights
after 200 training steps (tests/test_gpu_parity_big.py::test_gradient_parity_4096_rays_after_200_steps)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import robust_nerf_b200 as rn
from oracle import torch_ref as TR
import test_gpu_parity_big as T

dev = torch.device("cuda:0")
wc, wf = T._trained_weights(rn, dev)
B, seed = 4096, 8
ro, rd, _, _ = T._scene_rays(rn, dev, B, seed)
g = torch.Generator(device=dev).manual_seed(seed)
target = torch.rand(B, 3, device=dev, generator=g)
t_rand = torch.rand(B, 64, device=dev, generator=g)
u = torch.rand(B, 128, device=dev, generator=g)
nc, nf = T._net_from(rn, wc, dev), T._net_from(rn, wf, dev)
with torch.no_grad():
    out = rn.render_rays(nc, nf, ro, rd, rn.RenderConfig(), is_train=True, t_rand=t_rand, u=u, return_extras=True)
    pc, pf = TR.to_params(wc, dev, False), TR.to_params(wf, dev, False)
    res = TR.render_rays(pc, pf, ro, rd, is_train=True, t_rand=t_rand, u=u)
    for k in ("rgb_coarse", "rgb_fine"):
        d = (out[k] - res[k]).abs().max(-1)[0]
        print(k, "max", float(d.max()), "rays > 1e-2:", int((d > 1e-2).sum()), "> 1e-3:", int((d > 1e-3).sum()))
    d = (out["rgb_fine"] - res["rgb_fine"]).abs().max(-1)[0]
    bad = torch.nonzero(d > 1e-2).reshape(-1)[:5]
    for b in bad.tolist():
        print("ray", b, "rgb ours", out["rgb_fine"][b].tolist(), "ref", res["rgb_fine"][b].tolist())
        print("  coarse rgb diff", float((out["rgb_coarse"][b] - res["rgb_coarse"][b]).abs().max()))
        wd = (out["weights_coarse"][b] - res["weights_coarse"][b]).abs()
        print("  coarse weights max diff", float(wd.max()), "sum ours", float(out["weights_coarse"][b].sum()), "ref", float(res["weights_coarse"][b].sum()))
        zd = (out["z_fine"][b] - res["z_fine"][b]).abs()
        print("  z_fine max diff", float(zd.max()), "n>1e-4", int((zd > 1e-4).sum()))
        # evaluate the fine net on the SAME points (the reference's z) with both implementations
        z = res["z_fine"][b:b + 1]
        pts = ro[b:b + 1, None, :] + rd[b:b + 1, None, :] * z[..., None]
        vd = rd[b:b + 1] / rd[b:b + 1].norm(dim=-1, keepdim=True)
        raw = nf.forward_raw(pts.reshape(-1, 3).contiguous(), vd.contiguous(), z.shape[-1])
        rgb_t, sig_t = TR.nerf_forward(pf, pts.reshape(-1, 3), vd.expand(z.shape[-1], 3))
        print("  same points: sigma pre-act ours vs ref max diff", float((torch.relu(raw[:, 3]) - sig_t[:, 0]).abs().max()),
              "sigma max", float(sig_t.max()), "rgb diff", float((torch.sigmoid(raw[:, :3]) - rgb_t).abs().max()))
        wf_o, wf_r = out["weights_fine"][b], res["weights_fine"][b]
        i = int((wf_o - wf_r).abs().argmax())
        print("  fine weights max diff", float((wf_o - wf_r).abs().max()), "at sample", i, "z ours", float(out["z_fine"][b, i]), "ref", float(res["z_fine"][b, i]))
