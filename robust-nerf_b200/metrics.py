"""Evaluation metrics behind the reference's interface (noisy_src/metrics.py:15-116), computed on the GPU.

`compute_psnr`, `compute_mse`, `compute_ssim` keep the reference signatures (one (H,W,3) image pair -> 0-dim tensor);
`image_metrics` is the batched form the evaluation loop wants: every rendered view of a batch in ONE launch, results
stay on the device (no per-image `.item()` round trip).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from ._lib import call, lib, ptr, stream_ptr


def image_metrics(pred: torch.Tensor, target: torch.Tensor, max_val: float = 1.0, C1: float = 0.01 ** 2,
                  C2: float = 0.03 ** 2) -> Dict[str, torch.Tensor]:
    """pred, target: (N,H,W,3) fp32 CUDA tensors -> {"mse", "psnr", "ssim"}, each (N,) on the device."""
    p, t = ops._f32(pred, "pred"), ops._f32(target, "target")
    if p.dim() != 4 or p.shape[-1] != 3 or p.shape != t.shape:
        raise ValueError("image_metrics expects pred and target of shape (N, H, W, 3)")
    N, H, W, _ = p.shape
    mse = torch.empty(N, device=p.device, dtype=torch.float32)
    ssim = torch.empty(N, device=p.device, dtype=torch.float32)
    if N:
        scratch = torch.empty(lib().rn_image_metrics_scratch_bytes(N, H, W), device=p.device, dtype=torch.uint8)
        call("rn_image_metrics", ptr(p), ptr(t), N, H, W, float(C1), float(C2), ptr(scratch), ptr(mse), ptr(ssim), stream_ptr())
    # metrics.py:38-41 (inf for identical images falls out of log10(0))
    psnr = 20.0 * torch.log10(torch.tensor(float(max_val), device=p.device)) - 10.0 * torch.log10(mse)
    return {"mse": mse, "psnr": psnr, "ssim": ssim}


def _single(pred: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
    if pred.dim() != 3 or pred.shape[-1] != 3:
        raise NotImplementedError("only (H, W, 3) image pairs are implemented (the reference's evaluation path)")
    return image_metrics(pred.unsqueeze(0), target.unsqueeze(0))


def compute_mse(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return _single(pred, target)["mse"][0]


def compute_psnr(pred: torch.Tensor, target: torch.Tensor, max_val: float = 1.0) -> torch.Tensor:
    if pred.dim() != 3 or pred.shape[-1] != 3:
        raise NotImplementedError("only (H, W, 3) image pairs are implemented (the reference's evaluation path)")
    return image_metrics(pred.unsqueeze(0), target.unsqueeze(0), max_val=max_val)["psnr"][0]


def compute_ssim(pred: torch.Tensor, target: torch.Tensor, window_size: int = 11, C1: float = 0.01 ** 2,
                 C2: float = 0.03 ** 2) -> torch.Tensor:
    if window_size != 11:
        raise NotImplementedError("only the reference's default 11 x 11 window is implemented")
    if pred.dim() != 3 or pred.shape[-1] != 3:
        raise NotImplementedError("only (H, W, 3) image pairs are implemented (the reference's evaluation path)")
    return image_metrics(pred.unsqueeze(0), target.unsqueeze(0), C1=C1, C2=C2)["ssim"][0]
