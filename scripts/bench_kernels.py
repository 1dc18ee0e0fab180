#!/usr/bin/env python
"""Stand-alone bandwidth sweep of the HBM-bound kernels (BASELINE.json config 5: 65,536-ray batch,
128 coarse + 256 fine samples) plus config-4 style render throughput and the pose-opt training step.
Prints one JSON object; achieved GB/s uses the ALGORITHMIC bytes of SURVEY.md section 8(a):
  composite fwd 24*S+32, fwd+bwd 60*S+56 B per ray and pass; sample_pdf+merge 24*Nc+20*Nf+24; stratified 20*Nc+24.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn          # noqa: E402
from robust_nerf_b200 import ops       # noqa: E402


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()              # 512 MB write: evicts the 126 MB L2 between timed iterations
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e-3


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = float(peaks["hbm_gbs"])
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {"peak_gbs": peak, "kernels": {}}
    g = torch.Generator(device=dev).manual_seed(0)
    for (B, Nc, Nf) in ((65536, 128, 256), (4096, 64, 128)):
        tag = f"B{B}_{Nc}+{Nf}"
        ro = torch.randn(B, 3, device=dev, generator=g)
        rd = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g), dim=-1)
        zb = torch.linspace(2.0, 6.0, Nc, device=dev)
        t_rand = torch.rand(B, Nc, device=dev, generator=g)
        t = timeit(lambda: ops.stratified(ro, rd, zb, t_rand), flush=flush)
        out["kernels"][f"stratified_{tag}"] = {"us": t * 1e6, "GBs": B * (20 * Nc + 24) / t / 1e9, "frac": B * (20 * Nc + 24) / t / 1e9 / peak}
        z, _ = ops.stratified(ro, rd, zb, t_rand)
        w = torch.rand(B, Nc, device=dev, generator=g) ** 4
        u = torch.rand(B, Nf, device=dev, generator=g)
        t = timeit(lambda: ops.sample_hierarchical(ro, rd, z, w, u), flush=flush)
        by = B * (24 * Nc + 20 * Nf + 24)
        out["kernels"][f"sample_hierarchical_{tag}"] = {"us": t * 1e6, "GBs": by / t / 1e9, "frac": by / t / 1e9 / peak}
        for S in (Nc, Nc + Nf):
            raw = torch.randn(B, S, 4, device=dev, generator=g)
            raw[..., 3] = raw[..., 3].abs() * 10
            zz = torch.sort(torch.rand(B, S, device=dev, generator=g) * 4 + 2, -1)[0]
            outs = [torch.empty(B, 3, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev), torch.empty(B, S, device=dev)]
            d_raw = torch.empty(B, S, 4, device=dev)
            gm = torch.randn(B, 3, device=dev, generator=g)
            from robust_nerf_b200._lib import call, ptr, stream_ptr

            def fwd():
                call("rn_composite_fwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, 0.0, ptr(outs[0]), ptr(outs[1]),
                     ptr(outs[2]), ptr(outs[3]), stream_ptr())

            def bwd():
                call("rn_composite_bwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, ptr(gm), None, None, None, None,
                     None, ptr(d_raw), None, stream_ptr())
            tf, tb = timeit(fwd, flush=flush), timeit(bwd, flush=flush)
            bf, bb = B * (24 * S + 32), B * (36 * S + 24)
            out["kernels"][f"composite_fwd_{tag}_S{S}"] = {"us": tf * 1e6, "GBs": bf / tf / 1e9, "frac": bf / tf / 1e9 / peak}
            out["kernels"][f"composite_bwd_{tag}_S{S}"] = {"us": tb * 1e6, "GBs": bb / tb / 1e9, "frac": bb / tb / 1e9 / peak}
            out["kernels"][f"composite_fwd+bwd_{tag}_S{S}"] = {"us": (tf + tb) * 1e6, "GBs": (bf + bb) / (tf + tb) / 1e9,
                                                              "frac": (bf + bb) / (tf + tb) / 1e9 / peak}
    # ---- render throughput (config 4 shape: 800x800 views, 64+128, det sampling), one GPU ----
    torch.manual_seed(42)
    coarse, fine = rn.create_nerf()
    coarse, fine = coarse.to(dev), fine.to(dev)
    cfg = rn.RenderConfig()
    poses = rn.hemisphere_poses(2, seed=1, device=dev)
    focal = 0.5 * 800 / __import__("math").tan(0.5 * 0.6911112070083618)
    for tile in (32768, 131072):
        rn.render_views_sharded(coarse, fine, poses[:1], 800, 800, focal, cfg, tile_rays=tile)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = rn.render_views_sharded(coarse, fine, poses, 800, 800, focal, cfg, tile_rays=tile)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        out.setdefault("render", {})[f"tile{tile}"] = {"rays": res["rays_rendered"], "s": t, "Mrays_per_s": res["rays_rendered"] / t / 1e6,
                                                      "frac_of_mlp_roofline": res["rays_rendered"] / t * 303824896 / 1e12 / float(peaks.get("bf16_tflops_sustained", 1400.0))}
    # ---- joint pose-optimisation training step (config 3, one GPU shard: 4096 rays) ----
    scene = rn.make_scene(800, 800, 100, seed=0, device=dev)
    noisy = rn.add_noise_to_poses(scene.poses, 5.0, 5.0, seed=42)
    cam = rn.CameraPoseParameters(noisy).to(dev)
    with torch.no_grad():
        cam.rotation_deltas.normal_(0, 1e-3)
    ds, sampler = rn.create_pixel_dataset(scene)
    sampler.batch_size = 4096
    tr = rn.Trainer(coarse, fine, cfg, camera_params=cam, rotation_reg_weight=0.01, translation_reg_weight=0.001)
    batches = [sampler.sample_batch() for _ in range(4)]
    for i in range(5):
        tr.step_pixels(batches[i % 4], sampler)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        tr.step_pixels(batches[i % 4], sampler)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / 10
    out["pose_opt_step"] = {"ms": t * 1e3, "rays_per_s": 4096 / t}
    for i in range(5):
        tr.step_pixels_graphed(batches[i % 4], sampler)
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        tr.step_pixels_graphed(batches[i % 4], sampler)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / 20
    out["pose_opt_step_graphed"] = {"ms": t * 1e3, "rays_per_s": 4096 / t}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
