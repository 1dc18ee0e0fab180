"""Golden vectors for the evaluation metrics: run the UNMODIFIED reference
`noisy_src/metrics.py` (compute_psnr, compute_mse, compute_ssim) on seeded images in the authoring container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics.py      ->  tests/golden/metrics.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
from noisy_src.metrics import compute_psnr, compute_mse, compute_ssim  # noqa: E402

torch.set_num_threads(8)
rng = np.random.default_rng(123)
out = {}
for tag, (N, H, W) in {"small": (3, 37, 45), "tile": (2, 64, 96), "tiny": (2, 7, 9)}.items():
    target = rng.uniform(0, 1, (N, H, W, 3)).astype(np.float32)
    # predictions = smoothed target + noise, so SSIM is neither ~0 nor ~1
    pred = np.clip(0.7 * target + 0.3 * np.roll(target, 1, axis=2) + rng.normal(0, 0.05, target.shape), 0, 1).astype(np.float32)
    pred[0] = target[0] * 0.9 + 0.05                                   # a highly correlated pair
    psnr, mse, ssim = [], [], []
    for i in range(N):
        p, t = torch.from_numpy(pred[i]), torch.from_numpy(target[i])
        psnr.append(float(compute_psnr(p, t)))
        mse.append(float(compute_mse(p, t)))
        ssim.append(float(compute_ssim(p, t)))
    out[f"{tag}_pred"], out[f"{tag}_target"] = pred, target
    out[f"{tag}_psnr"], out[f"{tag}_mse"], out[f"{tag}_ssim"] = np.array(psnr, np.float64), np.array(mse, np.float64), np.array(ssim, np.float64)
    print(tag, psnr, ssim)
np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
