"""Worker of tests/test_gpu_multi.py: one process per GPU (torch.distributed.run, NCCL).  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import robust_nerf_b200 as rn

    B = 512                                              # rays per rank
    g = torch.Generator(device="cpu").manual_seed(5)
    poses = rn.lego_poses(dev)
    idx = torch.randint(0, 100 * 640000, (world * B,), generator=g).to(dev)
    img, uv, _ = rn.ops.pixel_gather(idx, 800, 800, None)
    ro, rd = rn.ops.RayGen.apply(img, uv, poses, 800, 800, rn.synthetic.focal_from_fov(800), 400.0, 400.0)
    tgt = torch.rand(world * B, 3, generator=g).to(dev)
    shard = lambda t, r: t[r * B:(r + 1) * B].contiguous()

    def make():
        torch.manual_seed(42)
        c, f = rn.create_nerf(rn.ModelConfig())
        return c.to(dev), f.to(dev)

    # ---- 1. one data-parallel gradient (early all-reduce of the fine slice + the rest) vs the same shards run locally ----
    c, f = make()
    tr = rn.Trainer(c, f, rn.RenderConfig(), lr=5e-4)
    assert tr.world == world and tr._early is not None
    torch.manual_seed(100 + rank)
    tr.step_rays(shard(ro, rank), shard(rd, rank), shard(tgt, rank), optimise=False)
    g_dp = tr.gflat.clone()                               # SUM over ranks (the mean's 1/world lives in the Adam kernel)
    c1, f1 = make()
    tr1 = rn.Trainer(c1, f1, rn.RenderConfig(), lr=5e-4)
    tr1.world, tr1._early = 1, None                       # single-process emulation of the same shards
    g_sum = torch.zeros_like(g_dp)
    for r in range(world):
        torch.manual_seed(100 + r)
        tr1.step_rays(shard(ro, r), shard(rd, r), shard(tgt, r), optimise=False)
        g_sum += tr1.gflat
    grad_diff = float((g_dp - g_sum).abs().max())
    grad_scale = float(g_sum.abs().max())
    nf = tr.net_offsets[1]
    diff_fine, diff_coarse = float((g_dp[:nf] - g_sum[:nf]).abs().max()), float((g_dp[nf:] - g_sum[nf:]).abs().max())
    # the same gradient with the early (side-stream) all-reduce switched off: one all-reduce of the whole buffer
    early = tr._early
    tr._early = None
    torch.manual_seed(100 + rank)
    tr.step_rays(shard(ro, rank), shard(rd, rank), shard(tgt, rank), optimise=False)
    diff_noearly = float((tr.gflat - g_sum).abs().max())
    tr._early = early
    # diagnostics of the early path
    diag = {}
    local = None
    for name in ("sync_before", "main_stream"):
        orig = tr._allreduce_early

        def hook(name=name):
            lo, hi = tr._early
            if name == "sync_before":
                torch.cuda.synchronize()
                orig()
            else:
                dist.all_reduce(tr.gflat[lo:hi], op=dist.ReduceOp.SUM)
                tr._early_issued = True
                tr._side.wait_stream(torch.cuda.current_stream())
        tr._allreduce_early = hook
        torch.manual_seed(100 + rank)
        tr.step_rays(shard(ro, rank), shard(rd, rank), shard(tgt, rank), optimise=False)
        d_ = (tr.gflat - g_sum).abs()
        diag[name] = float(d_.max())
        tr._allreduce_early = orig
    bad = (g_dp - g_sum).abs() > 0
    diag["n_bad"] = int(bad.sum()); diag["first_bad"] = int(torch.nonzero(bad)[0]) if bad.any() else -1
    diag["last_bad"] = int(torch.nonzero(bad)[-1]) if bad.any() else -1
    # ---- 2. optimised steps, eager and graph-replayed: replicas stay bit-identical ----
    for it in range(3):
        torch.manual_seed(200 + 10 * it + rank)
        tr.step_rays(shard(ro, rank), shard(rd, rank), shard(tgt, rank))
    for it in range(3):
        torch.manual_seed(300 + 10 * it + rank)
        tr.step_rays_graphed(shard(ro, rank), shard(rd, rank), shard(tgt, rank))
    torch.cuda.synchronize()
    ref = tr.flat.clone()
    dist.broadcast(ref, src=0)
    d = (tr.flat - ref).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    moved = float((tr.flat - tr1.flat).abs().max())       # tr1 never stepped: the parameters did move
    if rank == 0:
        print(json.dumps({"world": world, "grad_max_abs_diff_vs_local_sum": grad_diff, "grad_max_abs": grad_scale,
                          "diff_fine_slice": diff_fine, "diff_coarse_slice": diff_coarse, "diff_without_early_allreduce": diff_noearly, "diag": diag,
                          "replica_max_abs_diff": float(d.item()), "params_moved": moved}), flush=True)
    tr._graphs.clear()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
