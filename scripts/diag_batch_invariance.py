#!/usr/bin/env python
"""Are per-ray render outputs independent of the batch they are rendered in?  (full 4096-ray batch vs its halves)"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import robust_nerf_b200 as rn
from oracle import nerf_oracle as O
from test_gpu_parity import load_net
dev = torch.device("cuda:0")
mc, mf = load_net(rn, O.make_weights(41, sharpen=True), dev), load_net(rn, O.make_weights(42, sharpen=True), dev)
data = rn.make_scene(64, 64, 100, seed=3, device=dev)
ds, sampler = rn.create_pixel_dataset(data)
idx = torch.from_numpy(np.random.default_rng(77).integers(0, ds.n_pixels, 4096)).to(dev)
pb = sampler.batch_from_indices(idx)
with torch.no_grad():
    ro, rd = sampler.get_rays_for_batch(pb, data.poses)
cfg = rn.RenderConfig(perturb=False)
for train in (False, True):
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        full = rn.render_rays(mc, mf, ro, rd, cfg, is_train=train)
        h1 = rn.render_rays(mc, mf, ro[:2048].contiguous(), rd[:2048].contiguous(), cfg, is_train=train)
        h2 = rn.render_rays(mc, mf, ro[2048:].contiguous(), rd[2048:].contiguous(), cfg, is_train=train)
    for k in ("rgb_coarse", "rgb_fine", "depth_fine", "acc_fine"):
        cat = torch.cat([h1[k], h2[k]], 0)
        d = (full[k] - cat).abs()
        print("train" if train else "eval ", k, "max abs diff", float(d.max()), "rays differing", int((d.reshape(4096, -1).max(-1)[0] > 0).sum()))
