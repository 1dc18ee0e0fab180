"""Synthetic lego-shaped scene: 800x800, 100 train views on the real lego
camera layout, random images, random-init weights.  Host-side data preparation only."""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618          # standard Blender-lego field of view
_POSES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lego_train_poses.npy")


@dataclass
class BlenderData:
    """Same fields as the reference's container (noisy_src/data.py:25-47)."""
    images: torch.Tensor
    poses: torch.Tensor
    H: int
    W: int
    focal: float


def focal_from_fov(W: int, camera_angle_x: float = CAMERA_ANGLE_X) -> float:
    return 0.5 * W / math.tan(0.5 * camera_angle_x)      # noisy_src/data.py:150


def lego_poses(device="cpu") -> torch.Tensor:
    """The 100 lego train-split camera poses (recorded as ground_truth_poses in the reference's
    outputs/*/final_poses.pt; camera distance 4.0311)."""
    return torch.from_numpy(np.load(_POSES)).to(device)


def hemisphere_poses(n: int, radius: float = 4.0311, seed: int = 1, device="cpu") -> torch.Tensor:
    """Look-at-origin poses on the upper hemisphere (same frame construction as inference.py:345-357)."""
    rng = np.random.default_rng(seed)
    poses = []
    for _ in range(n):
        phi = rng.uniform(0, 2 * np.pi)
        cos_t = rng.uniform(0.1, 0.95)
        sin_t = np.sqrt(1 - cos_t * cos_t)
        pos = radius * np.array([sin_t * np.cos(phi), sin_t * np.sin(phi), cos_t])
        fwd = -pos / np.linalg.norm(pos)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, pos
        poses.append(c2w)
    return torch.from_numpy(np.stack(poses)).to(device)


def add_noise_to_poses(poses: torch.Tensor, rotation_noise_deg: float = 0.0, translation_noise_pct: float = 0.0,
                       seed: Optional[int] = None) -> torch.Tensor:
    """Noisy initial poses for a synthetic scene: shorthand for `noise.add_noise_to_poses` (noisy_src/noise.py:194-234
    semantics, CUDA arithmetic, the reference's draw order on the CPU generator) that returns only the poses."""
    from .noise import NoiseConfig, add_noise_to_poses as _add
    return _add(poses, NoiseConfig(rotation_noise_deg, 0.0, translation_noise_pct, seed))[0]


def make_scene(H: int = 800, W: int = 800, n_views: int = 100, seed: int = 0, device="cuda",
               poses: Optional[torch.Tensor] = None) -> BlenderData:
    """Random images torch.rand(n,H,W,3) from a seeded CPU generator + the lego camera layout."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n_views, H, W, 3, generator=g)
    if poses is None:
        poses = lego_poses() if n_views <= 100 else hemisphere_poses(n_views, seed=seed)
        poses = poses[:n_views]
    return BlenderData(images=images.to(device), poses=poses.to(device), H=H, W=W, focal=focal_from_fov(W))
