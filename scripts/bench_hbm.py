#!/usr/bin/env python
"""The `hbm_kernels` block of bench.py alone (BASELINE configs[4]); RN_B200_LIB selects the library build (A/B)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
r = bench.hbm_kernels_block(torch.device("cuda:0"), bench.measured_peaks())
print(os.environ.get("RN_B200_LIB", "default"), " ".join(f"{k}={v['us']:.0f}us/{v['frac']:.3f}" for k, v in r.items() if isinstance(v, dict)))
