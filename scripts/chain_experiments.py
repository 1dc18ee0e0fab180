#!/usr/bin/env python
"""Bottleneck experiments on the layer-chained forward kernel (inference, 131,072 rays x 256 points):
time the render with individual pipeline parts switched off (results are wrong by construction)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robust_nerf_b200 as rn
from robust_nerf_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(42)
coarse, fine = rn.create_nerf()
coarse, fine = coarse.to(dev), fine.to(dev)
M = 131072 * 192
pts = torch.randn(M, 3, device=dev)
dirs = torch.nn.functional.normalize(torch.randn(131072, 3, device=dev), dim=-1)
out = {}
names = {0: "all on", 1: "no B (weight) loads", 2: "no TMA stores", 4: "no A loads", 8: "no epilogue math",
         3: "no B loads, no stores", 7: "no loads, no stores", 15: "MMA + barriers only", 10: "no stores, no epilogue math",
         11: "no B, no stores, no epilogue", 31: "MMA only (no TMEM reads either)"}
pair_names = {0: "all on", 1: "no weight loads", 4: "no side-chunk loads", 8: "no epilogue math", 16: "no epilogue (no TMEM reads)",
              5: "no loads", 13: "no loads, no epilogue math", 21: "MMA + barriers only"}
import ctypes
def timed(fn, n=3):
    """kernel-only milliseconds of the GEMM launches (CUDA events around each launch, rn_prof_*), per call"""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.rn_prof_enable(1)
    for _ in range(n):
        fn()
    ms3, fl3, n3 = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_int * 4)()
    lib.rn_prof_collect(ms3, fl3, n3)
    lib.rn_prof_enable(0)
    return ms3[0] / n
lib.rn_set_flag(0, 2)
with torch.no_grad():
    for dbg in ([] if os.environ.get("RN_TRAIN_ONLY") == "1" else sorted(pair_names)):
        lib.rn_set_flag(1, dbg)
        ms = timed(lambda: fine.forward_raw(pts, dirs, 192))
        key = f"pair chain (inference) {pair_names[dbg]}"
        out[key] = {"ms": ms, "TFLOPs": M * 1186816 / (ms * 1e-3) / 1e12}
        print(f"{key:60s} {ms:8.2f} ms  {out[key]['TFLOPs']:7.1f} TFLOP/s", flush=True)
lib.rn_set_flag(1, 0)
# training-mode forward (stores + masks), smaller M to fit the activation workspace
Mt = 16384 * 192
pts_t, dirs_t = pts[:Mt].clone().requires_grad_(True), dirs[:16384]
for dbg, nm in ((0, "all on"), (1, "no weight loads"), (2, "no TMA stores"), (3, "no weight loads, no stores"),
                (8, "no epilogue math"), (10, "no stores, no epilogue math")):
    lib.rn_set_flag(1, dbg)
    ms = timed(lambda: fine.forward_raw(pts_t, dirs_t, 192))
    key = f"pair chain (training fwd, {Mt} pts) {nm}"
    out[key] = {"ms": ms, "TFLOPs": Mt * 1186816 / (ms * 1e-3) / 1e12}
    print(f"{key:60s} {ms:8.2f} ms  {out[key]['TFLOPs']:7.1f} TFLOP/s", flush=True)
lib.rn_set_flag(1, 0)
lib.rn_set_flag(0, 1)
ms = timed(lambda: fine.forward_raw(pts_t, dirs_t, 192))
print(f"{'L2 chain (training fwd) all on':60s} {ms:8.2f} ms  {Mt * 1186816 / (ms * 1e-3) / 1e12:7.1f} TFLOP/s", flush=True)
# the encode kernel alone (included in every number above)
from robust_nerf_b200 import ops as _ops
with torch.no_grad():
  for ring in (() if os.environ.get("RN_SKIP_OLD") == "1" else (0, 1)):
    lib.rn_set_flag(2, ring)
    for chain in ((0, 1) if ring == 0 else (1,)):
        lib.rn_set_flag(0, chain)
        for dbg in ([0] if chain == 0 else ([0, 1, 2, 15, 31] if ring else sorted(names))):
            lib.rn_set_flag(1, dbg)
            for _ in range(2):
                fine.forward_raw(pts, dirs, 192)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fine.forward_raw(pts, dirs, 192)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            key = f"ring={'(5,2)' if ring == 0 else '(3,3)'} chain={chain} {names[dbg]}"
            out[key] = {"ms": ms, "TFLOPs": M * 1186816 / (ms * 1e-3) / 1e12}
            print(f"{key:60s} {ms:8.2f} ms  {out[key]['TFLOPs']:7.1f} TFLOP/s", flush=True)
lib.rn_set_flag(1, 0); lib.rn_set_flag(0, 2); lib.rn_set_flag(2, 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "chain_exp.json"), "w"), indent=1)
