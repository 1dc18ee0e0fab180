#!/usr/bin/env python
"""BASELINE.json config 4: test-view rendering (800x800, 64+128 samples, deterministic sampling) sharded by ray tile over
the GPUs of one box, no collective on the data path.  One JSON line from rank 0: Mrays/s = all rays of all ranks / max time.

    python scripts/bench_render_sharded.py --views 8                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/bench_render_sharded.py --views 16
"""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--tile-rays", type=int, default=131072)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    coarse, fine = rn.create_nerf()
    coarse, fine = coarse.to(dev), fine.to(dev)
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    poses = rn.hemisphere_poses(args.views, seed=1, device=dev)
    cfg = rn.RenderConfig()
    out = torch.zeros(args.views, H * W, 3, device=dev)
    with torch.no_grad():
        rn.render_views_sharded(coarse, fine, poses[:world], H, W, focal, cfg, tile_rays=args.tile_rays, rank=rank, world=world, out=out)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = rn.render_views_sharded(coarse, fine, poses, H, W, focal, cfg, tile_rays=args.tile_rays, rank=rank, world=world, out=out)
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    rays = torch.tensor([float(res["rays_rendered"])], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    if rank == 0:
        total = float(rays.item())
        assert total == args.views * H * W, (total, args.views * H * W)
        mr = total / (float(ms.item()) * 1e-3) / 1e6
        print(json.dumps({"metric": "render Mrays/s, 800x800 views, 64+128 samples, tile-sharded, no collective", "value": mr,
                          "unit": "Mrays/s", "n_gpus": world, "views": args.views, "ms": float(ms.item()),
                          "frac_of_mlp_flop_roofline": mr / (world * 1394.4e12 / 303824896 / 1e6), "scaling": "strong"}), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
