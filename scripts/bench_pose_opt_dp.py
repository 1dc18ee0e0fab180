#!/usr/bin/env python
"""BASELINE.json configs[2]: joint pose optimisation (rotation + translation, 5 deg / 5 % noisy initial poses) as a
data-parallel training step -- every rank renders its own 4096-pixel shard of the batch, ONE all-reduce of the flat
gradient buffer (both MLPs + all pose parameters) per step, per-net clip 1.0 / pose clip 0.1, two Adam groups.

    python scripts/bench_pose_opt_dp.py                                             # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_pose_opt_dp.py                                               # N GPUs of one box

Prints one JSON line (rank 0): whole-job rays/s, ms/step (device time, max over ranks) and the check that the replicas
stayed identical (max |parameter difference| between rank 0 and every other rank, nets and poses: must be 0.0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RAYS_PER_GPU, STEPS, WARMUP, POOL = 4096, 20, 5, 4


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import robust_nerf_b200 as rn

    scene = rn.make_scene(800, 800, 100, seed=0, device=dev)
    noisy = rn.add_noise_to_poses(scene.poses, 5.0, 5.0, seed=42)          # same noisy initialisation on every rank
    cam = rn.CameraPoseParameters(noisy).to(dev)
    torch.manual_seed(42)
    coarse, fine = rn.create_nerf(rn.ModelConfig())
    coarse, fine = coarse.to(dev), fine.to(dev)
    cfg = rn.RenderConfig()
    tr = rn.Trainer(coarse, fine, cfg, camera_params=cam, rotation_reg_weight=0.01, translation_reg_weight=0.001)
    ds, sampler = rn.create_pixel_dataset(scene)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)            # every rank draws its own pixels
    batches = [sampler.batch_from_indices(torch.randint(0, ds.n_pixels, (RAYS_PER_GPU,), generator=g).to(dev))
               for _ in range(POOL)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(WARMUP):
        tr.step_pixels_graphed(batches[i % POOL], sampler)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS):
        loss = tr.step_pixels_graphed(batches[i % POOL], sampler)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    flat = tr.flat.detach().clone()
    diff = torch.zeros(1, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        diff = (flat - ref).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    err = cam.compute_pose_errors(scene.poses)
    if rank == 0:
        ms = float(t.item()) / STEPS
        print(json.dumps({
            "workload": "configs[2]: joint pose optimisation step (rotation + translation, 5 deg / 5 % noisy init), "
                        "4096 rays per GPU, 64+128 samples, data parallel, CUDA-graph replay",
            "n_gpus": world, "steps": STEPS, "warmup": WARMUP, "ms_per_step": ms,
            "rays_per_s": world * RAYS_PER_GPU / (ms * 1e-3), "loss": float(loss.item()),
            "replica_max_abs_param_diff": float(diff.item()),
            "pose_parameters_moved": float(torch.cat([p.detach().abs().reshape(-1) for p in cam.parameters()]).max().item()),
            "rotation_error_mean_deg": err["rotation_error_mean"], "translation_error_mean": err["translation_error_mean"]}),
              flush=True)
    if world > 1:
        tr._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
