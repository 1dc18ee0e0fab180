#!/usr/bin/env python
"""Two renders of 131,072 rays (64+128 samples) for an ncu launch list of the inference path:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/render_once.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn
dev = torch.device("cuda:0")
torch.manual_seed(42)
coarse, fine = rn.create_nerf(rn.ModelConfig())
coarse, fine = coarse.to(dev), fine.to(dev)
cfg = rn.RenderConfig()
scene_poses = rn.lego_poses(device=dev)
focal = 0.5 * 800 / torch.tan(torch.tensor(0.5 * 0.6911112070083618)).item()
with torch.no_grad():
    dirs = rn.get_ray_directions(800, 800, focal, device=dev).reshape(-1, 3)
    ro, rd = rn.get_rays(dirs[:131072], scene_poses[0])
    for _ in range(2):
        out = rn.render_rays(coarse, fine, ro, rd, cfg, is_train=False)
torch.cuda.synchronize()
print("ok", float(out["rgb_fine"].mean()))
