#!/usr/bin/env python
"""Benchmark of the render-and-train hot path (BASELINE.json metric: train rays/s, fwd+bwd,
800x800 lego-shaped synthetic scene).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port on host cores)

A step = one pass of the hot path over one batch: BASELINE.json configs[1], the clean-pose NeRF
training step on a 4096-ray batch per GPU (render 64 coarse + 128 fine samples, MSE coarse+fine,
backward, one all-reduce of the flat gradient buffer when N > 1, joint clip at 1.0, Adam).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
NC, NF = 64, 128
POINTS_PER_RAY = NC + (NC + NF)                     # coarse net + fine net evaluations
FLOP_PER_POINT_FWD = 1186816
TRAIN_FLOP_PER_RAY = 3 * FLOP_PER_POINT_FWD * POINTS_PER_RAY   # 911,474,688 (fwd + dgrad + wgrad)
METRIC = "train rays/s (fwd+bwd+clip+Adam), 4096-ray batch/GPU, 64+128 samples, 800x800 lego-shaped synthetic scene"
WORKLOAD = "configs[1]: clean-pose NeRF training step, 4096-ray batch, synthetic 800x800 lego-shaped scene (100 views)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d["bf16_tflops_sustained"]), "gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling.  The sampler is started before the warm-up (nvidia-smi needs ~100 ms
    to produce its first line, as long as a whole 20-step timed region) and the samples are then filtered to the
    wall-clock windows of the timed regions (`mark()` ... `mark()`)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/rn_clocks_{os.getpid()}.csv"
        self.windows, self._open = [], None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Open / close a timed-region window (call right before the first and right after the last synchronised step)."""
        now = time.time()
        if self._open is None:
            self._open = now
        else:
            self.windows.append((self._open, now))
            self._open = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        import datetime
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if any(a - 0.005 <= r[0] <= b + 0.005 for a, b in self.windows)]
        used = inside if inside else rows
        reasons = set()
        for r in used:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if used:
            out = {"sm_mhz": statistics.median([r[1] for r in used]), "sm_max_mhz": max(r[2] for r in used),
                   "reasons": sorted(reasons), "samples": len(used),
                   "window": "timed regions (value + e2e)" if inside else "whole run (no sample fell inside a timed region)"}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch path on the host cores
# ------------------------------------------------------------------------------------------------
REFERENCE_ROOT = "/root/reference"


def host_threads():
    """Threads the CPU arm may use: every core this process is allowed on.  Set explicitly -- torchrun exports
    OMP_NUM_THREADS=1, which would otherwise pin the N > 1 reference arms to one core."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n)


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class CpuArm:
    """Config 1 (one 100x100 view through the chunked renderer, eval mode) and the clean-pose training step (fwd + bwd +
    joint clip + Adam) in fp32 PyTorch on the host.  kind = "reference": the UNMODIFIED reference imported from
    /root/reference (authoring container only -- the tree does not exist on the GPU box); kind = "port":
    oracle/torch_ref.py, the same aten operators in the same order, pinned to the reference's golden vectors
    (tests/test_torch_ref_golden.py)."""

    def __init__(self):
        import numpy as np
        import torch
        self.torch, self.np = torch, np
        self.cores = host_threads()
        torch.set_num_threads(self.cores)
        self.kind = "port"
        self.poses = torch.from_numpy(np.load(os.path.join(ROOT, "robust-nerf_b200", "data", "lego_train_poses.npy"))).float()
        if os.path.isdir(os.path.join(REFERENCE_ROOT, "noisy_src")) and not os.environ.get("RN_BENCH_FORCE_PORT"):
            try:
                sys.dont_write_bytecode = True
                sys.path.insert(0, REFERENCE_ROOT)
                from noisy_src.config import ModelConfig, RenderConfig
                from noisy_src.model import create_nerf
                from noisy_src.rendering import NeRFRenderer
                from noisy_src import rays as R
                from noisy_src.train import train_step
                torch.manual_seed(42)
                c, f = create_nerf(ModelConfig())
                self.renderer = NeRFRenderer(c, f, RenderConfig())
                self.opt = torch.optim.Adam(self.renderer.parameters(), lr=5e-4)
                self._R, self._train_step = R, train_step
                self.kind = "reference"
            except Exception as e:      # noqa: BLE001  -- any import problem: fall back to the pinned port
                print(f"[bench] reference import failed ({e!r}); timing the torch port", file=sys.stderr)
        if self.kind == "port":
            from oracle import torch_ref as TR
            from oracle import nerf_oracle as O
            self.TR = TR
            self.pc, self.pf = TR.to_params(O.make_weights(21), "cpu"), TR.to_params(O.make_weights(22), "cpu")
            self.trainer = TR.RefTrainer(self.pc, self.pf, lr=5e-4)

    def _rays(self, n, seed):
        torch, np = self.torch, self.np
        H = W = 800
        focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
        g = torch.Generator().manual_seed(seed)
        img = torch.randint(0, 100, (n,), generator=g)
        u = torch.randint(0, W, (n,), generator=g).float()
        v = torch.randint(0, H, (n,), generator=g).float()
        d = torch.stack([(u - W / 2) / focal, -(v - H / 2) / focal, -torch.ones(n)], -1)
        Rm = self.poses[img][:, :3, :3]
        rd = torch.sum(d[:, None, :] * Rm, -1)
        rd = rd / rd.norm(dim=-1, keepdim=True)
        return self.poses[img][:, :3, 3].contiguous(), rd.contiguous(), torch.rand(n, 3, generator=g)

    def train_step(self, n_rays, seed):
        ro, rd, tgt = self._rays(n_rays, seed)
        t0 = time.perf_counter()
        if self.kind == "reference":
            self._train_step(self.renderer, self.opt, {"rays_o": ro, "rays_d": rd, "target_rgb": tgt})
        else:
            self.trainer.step_rays(ro, rd, tgt)
        return time.perf_counter() - t0

    def render_rays(self, n_rays):
        """The first n_rays rays of config 1's 100x100 view (pose 0, focal 138.889), chunk 4096, eval mode."""
        torch, np = self.torch, self.np
        H = W = 100
        focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                dirs = self._R.get_ray_directions(H, W, focal)
                ro, rd = self._R.get_rays(dirs, self.poses[0])
                self.renderer(ro.reshape(-1, 3)[:n_rays], rd.reshape(-1, 3)[:n_rays], chunk_size=4096, is_train=False)
            else:
                dirs = self.TR.get_ray_directions(H, W, focal)
                ro, rd = self.TR.get_rays(dirs, self.poses[0])
                ro, rd = ro.reshape(-1, 3)[:n_rays], rd.reshape(-1, 3)[:n_rays]
                for a in range(0, n_rays, 4096):
                    self.TR.render_rays(self.pc, self.pf, ro[a:a + 4096], rd[a:a + 4096], is_train=False)
        return time.perf_counter() - t0

    def describe(self):
        return {"cores": self.cores, "kind": self.kind, "cpu_model": _cpu_model(), "torch_threads": self.torch.get_num_threads()}


def cpu_baseline_block(train_rays=1024, render_rays=2048, train_steps=1):
    """Bounded CPU sample for the default run (about 10-30 s on 8 cores): one warm-up + `train_steps` timed 1024-ray
    training steps, and one pass over a slice of config 1's view."""
    arm = CpuArm()
    arm.train_step(train_rays, 0)
    ts = [arm.train_step(train_rays, 1 + i) for i in range(train_steps)]
    t_train = sum(ts) / len(ts)
    t_render = arm.render_rays(render_rays)
    d = arm.describe()
    return {"value": train_rays / t_train, "unit": "rays/s", "cores": d["cores"], "kind": d["kind"],
            "sample": f"{train_rays}-ray clean-pose training step (fwd + bwd + joint clip + Adam, fp32 PyTorch on the host, "
                      f"{train_steps} timed after 1 warm-up, {t_train:.2f} s/step); render: first {render_rays} rays of config 1's "
                      f"100x100 view, chunk 4096, eval mode, one pass ({t_render:.2f} s)",
            "render": {"value": render_rays / t_render, "unit": "rays/s", "config": "configs[0]: one 100x100 view, 64+128 samples, CPU"},
            "cpu_model": d["cpu_model"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_steps = max(1, args.steps) + max(0, args.warmup)
    # bounded sample: BASELINE.md section 3's 1024-ray step while the whole run stays within a few minutes (~5 s per
    # 1024 rays on 8 cores), smaller slices of the same step for long --steps
    sample = 1024 if n_steps <= 25 else int(max(128, (1024 * 25 // n_steps) // 64 * 64))
    if os.environ.get("RN_BENCH_CPU_RAYS"):                  # tests only: a tiny sample
        sample = int(os.environ["RN_BENCH_CPU_RAYS"])
    arm = CpuArm()
    for i in range(max(0, args.warmup)):
        arm.train_step(sample, i)
    times = [arm.train_step(sample, 100 + i) for i in range(max(1, args.steps))]
    t = sum(times) / len(times)
    n_render = 10000 if n_steps <= 25 else 2048
    if os.environ.get("RN_BENCH_CPU_RAYS"):
        n_render = 4 * int(os.environ["RN_BENCH_CPU_RAYS"])
    t_render = arm.render_rays(n_render)
    d = arm.describe()
    rate = sample / t
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_sampled": sample, "samples": f"{NC}+{NF}"},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": d["cores"], "kind": d["kind"], "cpu_model": d["cpu_model"],
                         "sample": f"{sample}-ray slice of the 4096-ray clean-pose training step per timed step (fp32 PyTorch on "
                                   f"the host: fwd + bwd + joint clip + Adam; "
                                   + ("the unmodified reference imported from /root/reference" if d["kind"] == "reference" else
                                      "oracle/torch_ref.py, the reference's aten operators in its order, pinned to its golden vectors")
                                   + f"); render: {n_render} rays of config 1's 100x100 view in {t_render:.2f} s",
                         "render": {"value": n_render / t_render, "unit": "rays/s",
                                    "config": "configs[0]: one 100x100 view, 64+128 samples, chunk 4096, eval mode"}},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# like-for-like GPU baseline: the fp32 eager PyTorch restatement on the same B200
# ------------------------------------------------------------------------------------------------
def torch_eager_numbers(dev, steps, warmup, rays=RAYS_PER_GPU, render_rays=131072):
    import torch
    from oracle import torch_ref as TR
    import robust_nerf_b200 as rn
    torch.backends.cuda.matmul.allow_tf32 = False           # the reference never enables TF32: plain fp32 cuBLAS
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(42)
    c, f = rn.create_nerf(rn.ModelConfig())
    strip = lambda m: {k: v for k, v in m.state_dict().items() if "freq_bands" not in k}
    pc, pf = TR.to_params(strip(c), dev), TR.to_params(strip(f), dev)
    tr = TR.RefTrainer(pc, pf, lr=5e-4)
    poses = rn.lego_poses(dev)
    dirs = TR.get_ray_directions(800, 800, TR.lego_focal(800), device=dev)
    g = torch.Generator(device="cpu").manual_seed(42)
    batches = []
    for _ in range(4):
        idx = torch.randint(0, 100 * 640000, (rays,), generator=g).to(dev)
        img, rem = idx // 640000, idx % 640000
        uv = torch.stack([(rem % 800).float(), (rem // 800).float()], -1)
        ro, rd = TR.rays_from_pixels(img, uv, poses, dirs)
        batches.append((ro.contiguous(), rd.contiguous(), torch.rand(rays, 3, device=dev)))
    for i in range(warmup):
        tr.step_rays(*batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = tr.step_rays(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ro, rd = TR.get_rays(dirs.reshape(-1, 3)[:render_rays], poses[0])
    with torch.no_grad():
        for _ in range(2):
            for a in range(0, render_rays, 32768):
                TR.render_rays(pc, pf, ro[a:a + 32768], rd[a:a + 32768], is_train=False)
        torch.cuda.synchronize()
        e0.record()
        for a in range(0, render_rays, 32768):
            TR.render_rays(pc, pf, ro[a:a + 32768], rd[a:a + 32768], is_train=False)
        e1.record()
        torch.cuda.synchronize()
    return {"train_rays_per_s": rays / (ms * 1e-3), "train_ms_per_step": ms, "loss": float(loss),
            "render_mrays_per_s": render_rays / (e0.elapsed_time(e1) * 1e-3) / 1e6,
            "what": "oracle/torch_ref.py (fp32 eager PyTorch restatement of the reference's path, autograd + torch.optim.Adam, "
                    "cuBLAS fp32 without TF32) on this GPU: 4096-ray clean-pose step; render of 131,072 rays in 32,768-ray chunks"}


def run_torch_eager(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    r = torch_eager_numbers(dev, max(1, args.steps), max(3, args.warmup))
    line = {"impl": "torch-eager", "metric": METRIC, "value": r["train_rays_per_s"], "unit": "rays/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["train_ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "samples": f"{NC}+{NF}", "note": r["what"]},
            "render": {"value": r["render_mrays_per_s"], "unit": "Mrays/s"}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
RENDER_FLOP_PER_RAY = FLOP_PER_POINT_FWD * POINTS_PER_RAY          # 303,824,896 (64 + 192 points, forward only)


def _timed(fn, barrier, reduce_max):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = fn()
    e1.record()
    barrier()
    return reduce_max(e0.elapsed_time(e1)), out


def render_block(rn, dev, rank, world, coarse, fine, cfg, focal, peaks, barrier, reduce_max, views_per_gpu):
    """BASELINE configs[3] shape on this job's GPUs: `views_per_gpu * world` test views of 800x800 (hemisphere poses, seed 1),
    64 + 128 samples, deterministic sampling, ray tiles dealt round-robin to the ranks, no collective.  value: poses and
    weights resident, images left on the device.  e2e: poses come from pinned host memory and every finished view is
    copied to pinned host memory inside the timed region (the copy of view v overlaps the render of view v+1)."""
    import torch
    H = W = 800
    n_views = views_per_gpu * world
    poses_host = rn.hemisphere_poses(n_views, seed=1).pin_memory()
    poses = poses_host.to(dev)
    out = torch.zeros(n_views, H * W, 3, device=dev)

    def run(p):
        with torch.no_grad():
            return rn.render_views_sharded(coarse, fine, p, H, W, focal, cfg, tile_rays=RENDER_TILE, rank=rank, world=world, out=out)

    run(poses[:world])                                       # warm-up: one view's worth of tiles per rank
    ms, res = _timed(lambda: run(poses), barrier, reduce_max)
    total_rays = n_views * H * W
    value = total_rays / (ms * 1e-3) / 1e6
    # e2e
    host_img = torch.empty(n_views, H * W, 3).pin_memory()
    copy_stream = torch.cuda.Stream()

    def run_e2e():
        p = poses_host.to(dev, non_blocking=True)
        with torch.no_grad():
            for v in range(n_views):
                rn.render_views_sharded(coarse, fine, p[v:v + 1], H, W, focal, cfg, tile_rays=RENDER_TILE, rank=rank, world=world,
                                        out=out[v:v + 1], view_offset=v)
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev)
                    host_img[v].copy_(out[v], non_blocking=True)
        copy_stream.synchronize()
        return None

    ms_e2e, _ = _timed(run_e2e, barrier, reduce_max)
    e2e_value = total_rays / (ms_e2e * 1e-3) / 1e6
    ceiling = world * peaks["tflops"] * 1e12 / RENDER_FLOP_PER_RAY / 1e6
    return {"workload": "configs[3] shape: 800x800 test views, 64+128 samples, deterministic sampling, tiles of "
                        f"{RENDER_TILE} rays dealt round-robin to the ranks, no collective",
            "views": n_views, "rays": total_rays, "value": value, "unit": "Mrays/s", "ms": ms,
            "roofline": {"bound": "tensor", "achieved": value * 1e6 * RENDER_FLOP_PER_RAY / 1e12, "peak": world * peaks["tflops"],
                         "unit": "TFLOP/s", "frac": value / ceiling, "ceiling_mrays_per_s": ceiling,
                         "algorithmic_flops_per_ray": RENDER_FLOP_PER_RAY},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms": ms_e2e, "h2d_bytes": n_views * 64,
                    "d2h_bytes": (n_views * H * W * 3 * 4) // world,
                    "note": "per rank: its tiles of every view device -> pinned host, overlapped with the next view"}}


RENDER_TILE = 131072


def pose_opt_block(rn, dev, rank, world, scene, cfg, barrier, reduce_max, peaks, steps=10, warmup=4):
    """BASELINE configs[2]: joint pose optimisation (rotation + translation, 5 deg / 5 % noisy initial poses, omega
    seeded N(0, 1e-3) so the rotation-gradient branch is live), 4096 pixels per GPU, ONE all-reduce of the flat gradient
    buffer (both MLPs + all pose parameters), per-net clip 1.0 / pose clip 0.1, two Adam groups, CUDA-graph replay."""
    import torch
    import torch.distributed as dist
    noisy = rn.add_noise_to_poses(scene.poses, 5.0, 5.0, seed=42)
    cam = rn.CameraPoseParameters(noisy).to(dev)
    with torch.no_grad():
        cam.rotation_deltas.copy_(torch.randn(100, 3, generator=torch.Generator().manual_seed(7)).to(dev) * 1e-3)
    torch.manual_seed(42)
    c, f = rn.create_nerf(rn.ModelConfig())
    c, f = c.to(dev), f.to(dev)
    tr = rn.Trainer(c, f, cfg, camera_params=cam, rotation_reg_weight=0.01, translation_reg_weight=0.001)
    ds, sampler = rn.create_pixel_dataset(scene)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    batches = [sampler.batch_from_indices(torch.randint(0, ds.n_pixels, (RAYS_PER_GPU,), generator=g).to(dev)) for _ in range(4)]
    for i in range(warmup):
        tr.step_pixels_graphed(batches[i % 4], sampler)

    def run():
        for i in range(steps):
            loss = tr.step_pixels_graphed(batches[i % 4], sampler)
        return loss

    ms, loss = _timed(run, barrier, reduce_max)
    ms /= steps
    flat = tr.flat.detach().clone()
    diff = 0.0
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        d = (flat - ref).abs().max().reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        diff = float(d.item())
    err = cam.compute_pose_errors(scene.poses)
    moved = float(torch.cat([p.detach().abs().reshape(-1) for p in cam.parameters()]).max().item())
    value = world * RAYS_PER_GPU / (ms * 1e-3)
    tr._graphs.clear()
    return {"workload": "configs[2]: joint pose optimisation step (rotation + translation, 5 deg / 5 % noisy init), 4096 pixels "
                        "per GPU, 64+128 samples, data parallel (one all-reduce of the flat gradient buffer incl. all pose "
                        "parameters), CUDA-graph replay",
            "value": value, "unit": "rays/s", "ms_per_step": ms, "steps": steps, "loss": float(loss.item()),
            "replica_max_abs_diff": diff, "pose_parameters_moved": moved,
            "rotation_error_mean_deg": err["rotation_error_mean"], "translation_error_mean": err["translation_error_mean"],
            "roofline": {"bound": "tensor", "achieved": value * TRAIN_FLOP_PER_RAY / 1e12, "peak": world * peaks["tflops"],
                         "unit": "TFLOP/s", "frac": value * TRAIN_FLOP_PER_RAY / 1e12 / (world * peaks["tflops"])}}


def hbm_kernels_block(dev, peaks, B=65536, Nc=128, Nf=256, iters=10, warmup=3):
    """BASELINE configs[4]: the HBM-bound kernels alone at 65,536 rays, 128 coarse + 256 fine samples.  achieved =
    ALGORITHMIC bytes / CUDA-event time on the launching stream; L2 flushed (512 MB write) between timed
    iterations; peak = the measured copy bandwidth."""
    import torch
    from robust_nerf_b200 import ops
    from robust_nerf_b200._lib import call, ptr, stream_ptr
    peak = peaks["gbs"]
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters * 1e-3

    def row(t, nbytes):
        return {"us": t * 1e6, "achieved": nbytes / t / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / t / 1e9 / peak,
                "algorithmic_bytes": nbytes}

    g = torch.Generator(device=dev).manual_seed(0)
    ro = torch.randn(B, 3, device=dev, generator=g)
    rd = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g), dim=-1)
    zb = torch.linspace(2.0, 6.0, Nc, device=dev)
    t_rand = torch.rand(B, Nc, device=dev, generator=g)
    out = {"workload": f"configs[4]: {B}-ray batch, {Nc} coarse + {Nf} fine samples, each kernel alone, L2 flushed between iterations",
           "bound": "hbm", "peak_source": peaks["source"]}
    out["stratified"] = row(timeit(lambda: ops.stratified(ro, rd, zb, t_rand)), B * (20 * Nc + 24))
    z, _ = ops.stratified(ro, rd, zb, t_rand)
    w = torch.rand(B, Nc, device=dev, generator=g) ** 4
    u = torch.rand(B, Nf, device=dev, generator=g)
    out["sample_pdf_merge"] = row(timeit(lambda: ops.sample_hierarchical(ro, rd, z, w, u)), B * (24 * Nc + 20 * Nf + 24))
    for S in (Nc, Nc + Nf):
        raw = torch.randn(B, S, 4, device=dev, generator=g)
        raw[..., 3] = raw[..., 3].abs() * 10                      # sigma ~ 10 |N(0,1)| (mirrors noisy_src/test_baseline.py:112-116)
        zz = torch.sort(torch.rand(B, S, device=dev, generator=g) * 4 + 2, -1)[0]
        outs = [torch.empty(B, 3, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev), torch.empty(B, S, device=dev)]
        d_raw = torch.empty(B, S, 4, device=dev)
        gm = torch.randn(B, 3, device=dev, generator=g)

        def fwd():
            call("rn_composite_fwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, 0.0, ptr(outs[0]), ptr(outs[1]),
                 ptr(outs[2]), ptr(outs[3]), stream_ptr())

        def bwd():
            call("rn_composite_bwd", None, None, ptr(raw), ptr(zz), ptr(rd), None, B, S, 1, ptr(gm), None, None, None, None,
                 None, ptr(d_raw), None, stream_ptr())
        tf, tb = timeit(fwd), timeit(bwd)
        bf, bb = B * (24 * S + 32), B * (36 * S + 24)
        out[f"composite_fwd_S{S}"] = row(tf, bf)
        out[f"composite_bwd_S{S}"] = row(tb, bb)
        out[f"composite_fwd_bwd_S{S}"] = row(tf + tb, bf + bb)
    del flush
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    # libraries (NCCL with NCCL_DEBUG=VERSION, ...) may print to stdout: keep fd 1 clean for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 through torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import robust_nerf_b200 as rn
    from robust_nerf_b200 import _lib
    lib = _lib.lib()

    # ---- synthetic scene: 100 views 800x800, random images, lego camera layout ----
    H = W = 800
    scene = rn.make_scene(H, W, 100, seed=0, device=dev)
    torch.manual_seed(42)                      # reference default seed (config.py:83)
    coarse, fine = rn.create_nerf(rn.ModelConfig())
    coarse, fine = coarse.to(dev), fine.to(dev)
    cfg = rn.RenderConfig()
    trainer = rn.Trainer(coarse, fine, cfg, lr=5e-4)
    ds, sampler = rn.create_pixel_dataset(scene)

    # clean-pose mode: batches of precomputed rays (RaySampler semantics, noisy_src/data.py:297-309)
    POOL = 8
    g = torch.Generator(device="cpu").manual_seed(42 + rank)
    dev_batches, host_batches = [], []
    for _ in range(POOL):
        idx = torch.randint(0, ds.n_pixels, (RAYS_PER_GPU,), generator=g).to(dev)
        pb = sampler.batch_from_indices(idx)
        with torch.no_grad():
            ro, rd = sampler.get_rays_for_batch(pb, scene.poses)
        dev_batches.append((ro.contiguous(), rd.contiguous(), pb.target_rgb.contiguous()))
        host_batches.append(tuple(t.cpu().pin_memory() for t in dev_batches[-1]))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W_, K_ = max(3, args.warmup), max(1, args.steps)

    use_graph = not args.no_graph
    step_fn = trainer.step_rays_graphed if use_graph else trainer.step_rays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- eager pass: K timed steps with every GEMM launch bracketed by CUDA events (roofline) and counted ----
    for i in range(W_):
        trainer.step_rays(*dev_batches[i % POOL])
    barrier()
    lib.rn_prof_enable(1)
    n0 = lib.rn_launch_count()
    e0.record()
    for i in range(K_):
        trainer.step_rays(*dev_batches[i % POOL])
    e1.record()
    barrier()
    launches = int(lib.rn_launch_count() - n0)
    lib.rn_prof_enable(0)
    ms3, fl3, ln3 = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_int * 4)()
    lib.rn_prof_collect(ms3, fl3, ln3)
    eager_ms_per_step = reduce_max(e0.elapsed_time(e1)) / K_

    # ---- device-resident timing (value): the same step, replayed from a CUDA graph unless --no-graph ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    def time_value():
        for i in range(W_):
            step_fn(*dev_batches[i % POOL])
        barrier()
        clocks.mark()
        e0.record()
        for i in range(K_):
            loss = step_fn(*dev_batches[i % POOL])
        e1.record()
        barrier()
        clocks.mark()
        return reduce_max(e0.elapsed_time(e1)), float(loss.item())

    # ---- end to end: pinned host batches -> device, step, loss back to the host, every step ----
    # An input pipeline as any trainer runs it: the upload of step i+1 is issued on a copy stream while step i runs, and
    # the host reads the loss of step i after it has LAUNCHED step i+1 (the D2H copy of every step's loss is enqueued
    # right behind that step; reading it one step late keeps the GPU fed instead of idling through a host round trip and
    # a graph launch per step).  Every step's H2D copy and D2H loss read happen inside the timed region; the loop ends
    # with the last loss on the host.
    loss_host = torch.zeros(2).pin_memory()
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            b = [t.to(dev, non_blocking=True) for t in host_batches[i % POOL]]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    def e2e_loop(n):
        nxt = upload(0)
        seen = []
        for i in range(n):
            b, ev = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in b:
                t.record_stream(cur)
            loss_host[i & 1:(i & 1) + 1].copy_(step_fn(*b).reshape(1), non_blocking=True)
            loss_events[i & 1].record(cur)
            if i + 1 < n:
                nxt = upload(i + 1)
            if i > 0:
                loss_events[(i - 1) & 1].synchronize()         # the caller reads the previous step's loss
                seen.append(float(loss_host[(i - 1) & 1]))
        loss_events[(n - 1) & 1].synchronize()
        seen.append(float(loss_host[(n - 1) & 1]))
        return seen

    def time_e2e():
        e2e_loop(max(3, W_))
        barrier()
        clocks.mark()
        e0.record()
        e2e_loop(K_)
        e1.record()
        barrier()
        clocks.mark()
        return reduce_max(e0.elapsed_time(e1))

    # The two timed regions run back to back on a power-capped part whose clocks sag as it warms; --e2e-first swaps
    # their order (a diagnostic: the later region is the slower one either way, profiles/r02_ab_log.md block 18).
    if args.e2e_first:
        t_e2e = time_e2e()
        t_ms, loss_val = time_value()
    else:
        t_ms, loss_val = time_value()
        t_e2e = time_e2e()
    clock_info = clocks.stop() if rank == 0 else None
    ms_per_step = t_ms / K_
    value = world * RAYS_PER_GPU * K_ / (t_ms * 1e-3)
    e2e_value = world * RAYS_PER_GPU * K_ / (t_e2e * 1e-3)

    # ---- the same step with the weight gradients AFTER the data-gradient chain (rn_set_flag(9, 0): 20 split-K launches per step
    # instead of the persistent stream beside the chain, csrc/wgrad_stream.cu) -- the A/B behind that default, same box ----
    seq_variant = None
    if world == 1 and not args.no_extras and use_graph:
        prev9 = ctypes.c_int(0)
        lib.rn_get_flag(9, ctypes.byref(prev9))
        if prev9.value > 0:
            try:
                trainer._graphs.clear()
                lib.rn_set_flag(9, 0)
                for i in range(W_):
                    step_fn(*dev_batches[i % POOL])
                torch.cuda.synchronize()
                e0.record()
                for i in range(K_):
                    step_fn(*dev_batches[i % POOL])
                e1.record()
                torch.cuda.synchronize()
                t_var = e0.elapsed_time(e1)
                seq_variant = {"flags": "9=0", "ms_per_step": t_var / K_, "value": RAYS_PER_GPU * K_ / (t_var * 1e-3), "unit": "rays/s",
                               "what": "weight gradients as 20 split-K launches after the data-gradient chain (timed after the "
                               "headline regions, i.e. on the warmer part)"}
            finally:
                lib.rn_set_flag(9, prev9.value)
                trainer._graphs.clear()

    # ---- strong scaling: the SAME global batch of 4096 rays split over the ranks ----
    strong = None
    if world > 1 and not args.no_extras and RAYS_PER_GPU % world == 0:
        rs = RAYS_PER_GPU // world
        small = [tuple(t[:rs].contiguous() for t in b) for b in dev_batches]
        for i in range(W_):
            step_fn(*small[i % POOL])
        barrier()
        e0.record()
        for i in range(K_):
            step_fn(*small[i % POOL])
        e1.record()
        barrier()
        t_strong = reduce_max(e0.elapsed_time(e1))
        strong = {"scaling": "strong", "global_batch_rays": RAYS_PER_GPU, "rays_per_gpu": rs, "ms_per_step": t_strong / K_,
                  "value": RAYS_PER_GPU * K_ / (t_strong * 1e-3), "unit": "rays/s",
                  "note": "same step, same global batch as the 1-GPU line; per-GPU work shrinks with N, the all-reduce does not"}

    # ---- what the collectives cost: the same K steps with the two all-reduces skipped ----
    allreduce = None
    if world > 1 and not args.no_extras:
        saved = [t.clone() for t in (trainer.flat, trainer.exp_avg, trainer.exp_avg_sq)]

        def timed_steps():
            trainer._graphs.clear()
            for i in range(W_):
                step_fn(*dev_batches[i % POOL])
            barrier()
            e0.record()
            for i in range(K_):
                step_fn(*dev_batches[i % POOL])
            e1.record()
            barrier()
            return reduce_max(e0.elapsed_time(e1)) / K_

        try:
            t_with = timed_steps()                       # back to back with the run below: same clocks
            trainer.skip_allreduce = True
            t_no = timed_steps()
        finally:
            trainer.skip_allreduce = False
            trainer._graphs.clear()
            for t, sv in zip((trainer.flat, trainer.exp_avg, trainer.exp_avg_sq), saved):
                t.copy_(sv)                     # the replicas diverged while the gradients were not averaged
            for m in trainer.nets:
                m._packed.key = None
        nbytes = 4 * int(trainer.gflat.numel())
        allreduce = {"collective": "ncclAllReduce(SUM) of the flat fp32 gradient buffer in two pieces: the fine network's slice on a "
                                   "side stream as soon as its backward has finished, the rest after the coarse backward",
                     "bytes_per_step": nbytes, "ms_per_step_with": t_with, "ms_per_step_without": t_no,
                     "exposed_us_per_step": (t_with - t_no) * 1e3,
                     "note": "exposed = step time with the collectives - step time without (same run, max over ranks); it "
                             "includes waiting for the slowest rank"}

    peaks = measured_peaks()
    trainer._graphs.clear()                       # release the captured step (and its 11 GB workspace) before the other blocks
    step_fn = None
    del trainer
    torch.cuda.empty_cache()

    # ---- render block (BASELINE configs[3] shape: 800x800 test views, 64+128, deterministic sampling, tile-sharded) ----
    render = None
    if not args.no_extras:
        render = render_block(rn, dev, rank, world, coarse, fine, cfg, scene.focal, peaks, barrier, reduce_max, args.render_views)
    # ---- pose-opt block (BASELINE configs[2]: joint pose optimisation step, data parallel at this N) ----
    pose_opt = None
    if not args.no_extras:
        pose_opt = pose_opt_block(rn, dev, rank, world, scene, cfg, barrier, reduce_max, peaks)
    # ---- HBM-bound kernels at BASELINE configs[4] (65,536 rays, 128 + 256 samples), rank 0 ----
    hbm = None
    if rank == 0 and not args.no_extras:
        hbm = hbm_kernels_block(dev, peaks)
    # ---- like-for-like: fp32 eager PyTorch on this GPU (rank 0, single-GPU runs) ----
    eager = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            eager = torch_eager_numbers(dev, 5, 3)
        except Exception as e:      # noqa: BLE001
            eager = {"error": repr(e)}
    extra = {}
    if seq_variant:
        extra["wgrad_sequential_variant"] = seq_variant
    if strong:
        extra["strong_scaling"] = strong
    if allreduce:
        extra["allreduce"] = allreduce
    if world > 1:
        dist.barrier()

    if rank == 0:
        gemm_ms_per_step = (ms3[0] + ms3[1] + ms3[2] + ms3[3]) / K_
        algo_flops_per_step = TRAIN_FLOP_PER_RAY * RAYS_PER_GPU
        achieved = algo_flops_per_step / (gemm_ms_per_step * 1e-3) / 1e12 if gemm_ms_per_step > 0 else None
        n_gemm = ln3[0] + ln3[1] + ln3[2] + ln3[3]
        traffic, hbm_view = None, None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                # second view of the same launches: the present data flow (activations and their gradients written
                # once and read once by the weight-gradient GEMMs) is HBM-heavy; DRAM bytes from the committed ncu pass
                # over the live GEMM time of this run
                step_bytes = tj.get("dram_bytes_per_step")
                if step_bytes and gemm_ms_per_step > 0:
                    gbs = step_bytes / (gemm_ms_per_step * 1e-3) / 1e9
                    hbm_view = {"dram_bytes_per_step": step_bytes, "achieved": gbs, "peak": peaks["gbs"], "unit": "GB/s",
                                "frac": gbs / peaks["gbs"], "source": "profiles/gemm_traffic.json (ncu) / live GEMM time"}
            except Exception:
                traffic, hbm_view = None, None
        per_mode = {}
        for i, nm in enumerate(("nt_forward", "nn_dgrad", "tn_wgrad", "dgrad_beside_wgrad")):
            if ln3[i]:
                per_mode[nm] = {"launches_per_step": ln3[i] / K_, "ms_per_step": ms3[i] / K_,
                                "executed_tflops": fl3[i] / (ms3[i] * 1e-3) / 1e12}
        cpu_block = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_block = cpu_baseline_block()
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "global_batch_rays": world * RAYS_PER_GPU,
                       "samples": f"{NC}+{NF}", "mlp": "8x256, PE L=10/4, bf16 tcgen05 operands, fp32 accumulate",
                       "parallelism": f"dp{world} (one all-reduce of the flat gradient buffer per step)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: 11 GB of activations and activation gradients streamed per step vs 126 MB L2, no flush needed",
                       "launch": "CUDA graph replay of the whole step" if use_graph else "eager (one launch per kernel)",
                       "eager_ms_per_step": eager_ms_per_step,
                       "loss": loss_val},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": RAYS_PER_GPU * 9 * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": t_e2e / K_},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMM kernels of the MLP: rn::mlp_chain_pair_kernel (forward chain), rn::mlp_chain_pair_bwd_kernel (data-gradient chain) with rn::wgrad_stream_kernel (weight gradients) beside it",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": traffic,
                         "peak_source": peaks["source"],
                         "algorithmic_flops_per_step": algo_flops_per_step, "gemm_launches_per_step": n_gemm / K_,
                         "gemm_ms_per_step": gemm_ms_per_step, "gemm_share_of_step": gemm_ms_per_step / eager_ms_per_step,
                         "measured_in": "eager pass of the same K steps (CUDA events around every GEMM launch on its stream)",
                         "per_mode": per_mode, "hbm_view": hbm_view},
            "cpu_baseline": cpu_block,
            "render": render, "pose_opt": pose_opt, "hbm_kernels": hbm, "torch_eager_same_gpu": eager,
            "extra": extra,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        # Orderly tear-down: every captured graph (they hold NCCL kernels) is gone, the device is drained and the ranks
        # have met, THEN the process group is destroyed and the interpreter exits normally.
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--render-views", type=int, default=2, help="800x800 test views per GPU in the render block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--e2e-first", action="store_true", help="time the end-to-end region before the device-resident one")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
