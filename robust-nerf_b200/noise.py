"""Camera-pose noise initialisation and pose-error tracking.

Mirror of the reference's `noisy_src/noise.py` interface (NoiseConfig :18-59, set_noise_seed :62-65,
add_noise_to_poses :194-234, compute_pose_error :237-268) with the per-pose Python loops replaced by one CUDA launch
(`rn_pose_noise`, `rn_pose_errors`; csrc/pose_noise.cu) and one device->host copy instead of three `.item()` per pose.

Random numbers: the reference draws, per pose and on the pose tensor's device, `randn(1)` and `randn(3)` for the
rotation and `randn(3)` for the translation.  Here the draws always come from torch's CPU generator in exactly that
order (`draw_pose_noise`), so a seeded run reproduces the reference's CPU-side initialisation (its data loader hands
CPU poses to `add_noise_to_poses`); the arithmetic runs on the GPU.  There is no CPU implementation of the arithmetic
in the package: without the CUDA library these functions raise.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from ._lib import call, lib, ptr, stream_ptr


@dataclass
class NoiseConfig:
    """Configuration of the pose noise (noise.py:18-59)."""

    rotation_noise_deg: float = 0.0      # rotation noise std in degrees
    translation_noise: float = 0.0       # translation noise std (scene units)
    translation_noise_pct: float = 0.0   # translation noise std (percent of the camera distance)
    seed: Optional[int] = None

    def __str__(self) -> str:
        parts = []
        if self.rotation_noise_deg > 0:
            parts.append(f"rot{self.rotation_noise_deg:.1f}deg")
        if self.translation_noise_pct > 0:
            parts.append(f"trans{self.translation_noise_pct:.1f}pct")
        elif self.translation_noise > 0:
            parts.append(f"trans{self.translation_noise:.3f}")
        return "_".join(parts) if parts else "clean"

    @property
    def has_noise(self) -> bool:
        return self.rotation_noise_deg > 0 or self.translation_noise > 0 or self.translation_noise_pct > 0

    def get_translation_std(self, camera_distance: float) -> float:
        if self.translation_noise_pct > 0:
            return camera_distance * (self.translation_noise_pct / 100.0)
        return self.translation_noise


def set_noise_seed(seed: int) -> None:
    """noise.py:62-65."""
    torch.manual_seed(seed)
    np.random.seed(seed)


def draw_pose_noise(poses: torch.Tensor, noise_config: NoiseConfig
                    ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """The standard-normal draws of add_noise_to_poses in the reference's order (host logic, CPU generator):
    per pose `randn(1)`, `randn(3)` when rotation noise is on, then `randn(3)` when that pose's translation std is
    positive (noise.py:96-101, 135, 176-187).  Returns (g_angle [n], g_axis [n,3], g_trans [n,3]) or None each."""
    if noise_config.seed is not None:
        set_noise_seed(noise_config.seed)
    n = poses.shape[0]
    rot = noise_config.rotation_noise_deg > 0
    tra = noise_config.translation_noise_pct > 0 or noise_config.translation_noise > 0
    dist = torch.norm(poses.detach()[:, :3, 3].float().cpu(), dim=-1).tolist() if tra else None
    g_angle = torch.zeros(n) if rot else None
    g_axis = torch.zeros(n, 3) if rot else None
    g_trans = torch.zeros(n, 3) if tra else None
    for i in range(n):
        if rot:
            g_angle[i] = torch.randn(1)[0]
            g_axis[i] = torch.randn(3)
        if tra and noise_config.get_translation_std(dist[i]) > 0:
            g_trans[i] = torch.randn(3)
    return g_angle, g_axis, g_trans


def _device_for(t: torch.Tensor) -> torch.device:
    lib()                                            # raises if the CUDA library is not built: no CPU path
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("robust_nerf_b200.noise needs a CUDA device (poses may live on the CPU, the arithmetic does not)")
    return torch.device("cuda", torch.cuda.current_device())


def add_noise_to_poses(poses: torch.Tensor, noise_config: NoiseConfig) -> Tuple[torch.Tensor, List[dict]]:
    """noise.py:194-234: (N,4,4) poses -> (noisy poses on the same device, per-pose noise_info dicts)."""
    g_angle, g_axis, g_trans = draw_pose_noise(poses, noise_config)
    dev = _device_for(poses)
    n = int(poses.shape[0])
    P = poses.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(P)
    info = torch.zeros(n, 2, device=dev, dtype=torch.float32)
    up = [None if g is None else g.to(dev).contiguous() for g in (g_angle, g_axis, g_trans)]
    std_rad = noise_config.rotation_noise_deg * np.pi / 180.0
    with torch.cuda.device(dev):
        call("rn_pose_noise", ptr(P), n, ptr(up[0]), ptr(up[1]), ptr(up[2]), float(std_rad), float(noise_config.translation_noise),
             float(noise_config.translation_noise_pct), ptr(out), ptr(info), stream_ptr())
    info_h = info.cpu().numpy()                      # the one host synchronisation
    dist = torch.norm(poses.detach()[:, :3, 3].float().cpu(), dim=-1).tolist()
    info_list = []
    for i in range(n):
        d = {"rotation_noise_deg": noise_config.rotation_noise_deg,
             "translation_noise": noise_config.get_translation_std(dist[i])}
        if noise_config.rotation_noise_deg > 0:
            d["actual_rotation_deg"] = float(info_h[i, 0])
        if d["translation_noise"] > 0:
            d["actual_translation_norm"] = float(info_h[i, 1])
        info_list.append(d)
    return out.to(poses.device), info_list


def compute_pose_errors_batch(poses_gt: torch.Tensor, poses_est: torch.Tensor) -> torch.Tensor:
    """[n,2] (rotation_error_deg, translation_error) of compute_pose_error for n pose pairs, on the GPU."""
    dev = _device_for(poses_est)
    G = poses_gt.detach().to(device=dev, dtype=torch.float32).contiguous()
    C = poses_est.detach().to(device=dev, dtype=torch.float32).contiguous()
    if G.shape != C.shape or G.dim() != 3 or tuple(G.shape[1:]) != (4, 4):
        raise ValueError(f"pose batches must both be (n,4,4), got {tuple(G.shape)} and {tuple(C.shape)}")
    err = torch.empty(G.shape[0], 2, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        call("rn_pose_errors", ptr(G), ptr(C), int(G.shape[0]), ptr(err), stream_ptr())
    return err


def compute_pose_error(pose_gt: torch.Tensor, pose_noisy: torch.Tensor) -> dict:
    """noise.py:237-268 for one pair of 4x4 poses."""
    e = compute_pose_errors_batch(pose_gt[None], pose_noisy[None]).cpu().numpy()
    return {"rotation_error_deg": float(e[0, 0]), "translation_error": float(e[0, 1])}
