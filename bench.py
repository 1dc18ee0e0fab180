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
# CPU arm: the oracle port timed on host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_rate(sample_rays: int, steps: int, warmup: int):
    """rays/s of the numpy oracle's full training step (render fwd, loss, bwd, joint clip, Adam) on a
    bounded sample of the 4096-ray batch."""
    import numpy as np
    from oracle import nerf_oracle as O
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    wc, wf = O.make_weights(21), O.make_weights(22)
    poses = np.load(os.path.join(ROOT, "robust-nerf_b200", "data", "lego_train_poses.npy"))
    H = W = 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    state, times = {}, []
    for it in range(warmup + steps):
        img = rng.integers(0, 100, sample_rays)
        uv = np.stack([rng.integers(0, W, sample_rays), rng.integers(0, H, sample_rays)], -1).astype(np.float32)
        ro, rd = O.get_rays_from_pixels(img, uv, poses, H, W, focal)
        target = rng.uniform(0, 1, (sample_rays, 3)).astype(np.float32)
        t_rand = rng.uniform(0, 1, (sample_rays, NC)).astype(np.float32)
        u = rng.uniform(0, 1, (sample_rays, NF)).astype(np.float32)
        t0 = time.perf_counter()
        out = O.train_step_grads(wc, wf, ro, rd, target, t_rand=t_rand, u=u)
        O.clip_and_adam([wc, wf], [out["grads_coarse"], out["grads_fine"]], state)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    return sample_rays / mean_t, cores, mean_t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: keep the whole --steps/--warmup run to a few minutes of CPU time (~5.5 s per 512 rays on 8 cores)
    n_steps = max(1, args.steps) + max(0, args.warmup)
    sample = int(min(512, max(64, (512 * 20 // n_steps) // 64 * 64)))
    rate, cores, t = cpu_train_step_rate(sample, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_sampled": sample, "samples": f"{NC}+{NF}"},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}-ray slice of the 4096-ray training step per timed step (numpy oracle port of the "
                                   "reference's PyTorch path: fwd + bwd + joint clip + Adam)"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
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
    ms3, fl3, ln3 = (ctypes.c_double * 3)(), (ctypes.c_double * 3)(), (ctypes.c_int * 3)()
    lib.rn_prof_collect(ms3, fl3, ln3)
    eager_ms_per_step = reduce_max(e0.elapsed_time(e1)) / K_

    # ---- device-resident timing (value): the same step, replayed from a CUDA graph unless --no-graph ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
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
    t_ms = reduce_max(e0.elapsed_time(e1))
    ms_per_step = t_ms / K_
    value = world * RAYS_PER_GPU * K_ / (t_ms * 1e-3)
    loss_val = float(loss.item())

    # ---- end to end: pinned host batches -> device, step, loss back to the host, every step ----
    # The upload of step i+1 is issued on a copy stream while step i runs (double buffering, as any input pipeline does);
    # every step's H2D copy and D2H loss read still happen inside the timed region, and the host waits for each loss.
    loss_host = torch.zeros(1).pin_memory()
    copy_stream = torch.cuda.Stream()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            b = [t.to(dev, non_blocking=True) for t in host_batches[i % POOL]]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    def e2e_loop(n):
        nxt = upload(0)
        for i in range(n):
            b, ev = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in b:
                t.record_stream(cur)
            loss_host.copy_(step_fn(*b).reshape(1), non_blocking=True)
            if i + 1 < n:
                nxt = upload(i + 1)
            cur.synchronize()                              # the caller reads the step's loss
            _ = float(loss_host[0])

    e2e_loop(3)
    barrier()
    clocks.mark()
    e0.record()
    e2e_loop(K_)
    e1.record()
    barrier()
    clocks.mark()
    clock_info = clocks.stop() if rank == 0 else None
    t_e2e = reduce_max(e0.elapsed_time(e1))
    e2e_value = world * RAYS_PER_GPU * K_ / (t_e2e * 1e-3)

    # ---- extras (not part of the contract line's headline): render throughput, pose-opt step ----
    extra = {}
    if rank == 0 and not args.no_extras:
        with torch.no_grad():
            dirs = rn.get_ray_directions(H, W, scene.focal, device=dev).reshape(-1, 3)
            ro, rd = rn.get_rays(dirs[:131072], scene.poses[0])
            for _ in range(2):
                rn.render_rays(coarse, fine, ro, rd, cfg, is_train=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                rn.render_rays(coarse, fine, ro, rd, cfg, is_train=False)
            e1.record()
            torch.cuda.synchronize()
            extra["render_mrays_per_s_1gpu"] = 3 * 131072 / (e0.elapsed_time(e1) * 1e-3) / 1e6
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks = measured_peaks()
        gemm_ms_per_step = (ms3[0] + ms3[1] + ms3[2]) / K_
        algo_flops_per_step = TRAIN_FLOP_PER_RAY * RAYS_PER_GPU
        achieved = algo_flops_per_step / (gemm_ms_per_step * 1e-3) / 1e12 if gemm_ms_per_step > 0 else None
        n_gemm = ln3[0] + ln3[1] + ln3[2]
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
        for i, nm in enumerate(("nt_forward", "nn_dgrad", "tn_wgrad")):
            if ln3[i]:
                per_mode[nm] = {"launches_per_step": ln3[i] / K_, "ms_per_step": ms3[i] / K_,
                                "executed_tflops": fl3[i] / (ms3[i] * 1e-3) / 1e12}
        cpu_rate, cores, cpu_t = (None, None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu_rate, cores, cpu_t = cpu_train_step_rate(256, 2, 1)
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
            "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMM kernels of the MLP: rn::mlp_chain_pair_kernel (forward chain), rn::mlp_chain_pair_bwd_kernel (data-gradient chain), rn::gemm_kernel<BN,2> (split-K weight gradients)",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None, "traffic": traffic,
                         "peak_source": peaks["source"],
                         "algorithmic_flops_per_step": algo_flops_per_step, "gemm_launches_per_step": n_gemm / K_,
                         "gemm_ms_per_step": gemm_ms_per_step, "gemm_share_of_step": gemm_ms_per_step / eager_ms_per_step,
                         "measured_in": "eager pass of the same K steps (CUDA events around every GEMM launch on its stream)",
                         "per_mode": per_mode, "hbm_view": hbm_view},
            "cpu_baseline": None if cpu_rate is None else {
                "value": cpu_rate, "unit": "rays/s", "cores": cores, "kind": "port",
                "sample": f"256-ray slice of the same training step, 2 timed steps after 1 warm-up ({cpu_t:.1f} s/step), numpy oracle port"},
            "extra": extra,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        # Tear-down with captured NCCL kernels alive can block in destroy_process_group: release the graphs,
        # drain the device, meet once more, then leave without running NCCL's destructors.
        trainer._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
