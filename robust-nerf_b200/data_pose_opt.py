"""Pixel batches for joint pose optimisation behind the reference's interface
(noisy_src/data_pose_opt.py:21-244).

The reference materialises three tables over all N*H*W pixels (int64 image index, float32 (u,v),
rgb: 1.8 GB at 100 x 800^2).  Here the bookkeeping is index arithmetic inside the gather kernel
(image = idx // (H*W), v = (idx % (H*W)) // W, u = idx % W) -- bit-identical values, no tables --
and ray generation is one fused launch instead of a Python loop over unique images.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Tuple

import torch

from . import ops


@dataclass
class PixelBatch:
    image_indices: torch.Tensor  # (B,) int64
    pixel_coords: torch.Tensor   # (B, 2) float32 (u, v)
    target_rgb: torch.Tensor     # (B, 3)


class PixelDataset:
    def __init__(self, data: Any):
        """`data` is anything with .images (N,H,W,3), .poses, .H, .W, .focal (the reference's BlenderData)."""
        self.H, self.W, self.focal = int(data.H), int(data.W), float(data.focal)
        self.device = data.images.device
        if self.device.type != "cuda":
            raise RuntimeError("PixelDataset needs CUDA-resident images (no CPU fallback)")
        # fp32 (N,H,W,3) as the reference keeps them, or uint8 (N,H,W,3): the reference's images are exact multiples of
        # 1/255 (data.py:118-136), so a uint8 table with the /255 done inside the gather kernel is lossless and a quarter
        # of the size (192 MB instead of 768 MB at 100 x 800^2); `ops.quantize_images` converts and checks.
        self.images = data.images.contiguous()
        self.n_images = self.images.shape[0]
        self.n_pixels = self.n_images * self.H * self.W
        self._ray_directions = None

    @property
    def target_rgb(self) -> torch.Tensor:
        """(N*H*W, 3) fp32 table of the reference (data_pose_opt.py:76): a view for fp32 storage, materialised on demand
        for uint8 storage."""
        if self.images.dtype == torch.uint8:
            return ops.dequantize_images(self.images).reshape(-1, 3)
        return self.images.reshape(-1, 3)

    @property
    def ray_directions(self) -> torch.Tensor:
        if self._ray_directions is None:
            self._ray_directions = ops.ray_directions(self.H, self.W, self.focal, self.W / 2.0, self.H / 2.0, self.device)
        return self._ray_directions

    # the reference's tables, materialised only if somebody asks for them
    @property
    def image_indices(self) -> torch.Tensor:
        return torch.arange(self.n_images, device=self.device).repeat_interleave(self.H * self.W)

    @property
    def pixel_coords(self) -> torch.Tensor:
        flat = torch.arange(self.H * self.W, device=self.device)
        uv = torch.stack([(flat % self.W).float(), (flat // self.W).float()], -1)
        return uv.repeat(self.n_images, 1)

    def _raygen(self, image_indices, pixel_coords, poses):
        return ops.RayGen.apply(image_indices, pixel_coords, poses, self.H, self.W, self.focal, self.W / 2.0, self.H / 2.0)

    def get_rays_from_pixels(self, pixel_batch: PixelBatch, poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """`poses` holds one pose per UNIQUE image of the batch, ordered by image index
        (data_pose_opt.py:105-122); rays come back in batch order."""
        _, rank = torch.unique(pixel_batch.image_indices, return_inverse=True)
        return self._raygen(rank, pixel_batch.pixel_coords, poses)


class PixelSampler:
    def __init__(self, dataset: PixelDataset, batch_size: int = 1024):
        self.dataset = dataset
        self.batch_size = batch_size
        self.device = dataset.device
        self.n_pixels = dataset.n_pixels

    def sample_batch(self) -> PixelBatch:
        """torch.randint with replacement (data_pose_opt.py:188-192), then one gather kernel."""
        indices = torch.randint(0, self.n_pixels, (self.batch_size,), device=self.device)
        return self.batch_from_indices(indices)

    def batch_from_indices(self, indices: torch.Tensor) -> PixelBatch:
        img, uv, rgb = ops.pixel_gather(indices, self.dataset.H, self.dataset.W, self.dataset.images)
        return PixelBatch(image_indices=img, pixel_coords=uv, target_rgb=rgb)

    def get_rays_for_batch(self, pixel_batch: PixelBatch, poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """poses: all (N_images,4,4) current poses (data_pose_opt.py:200-223); net effect poses[image_idx]."""
        return self.dataset._raygen(pixel_batch.image_indices, pixel_batch.pixel_coords, poses)

    def get_rays_for_batch_fused(self, pixel_batch: PixelBatch, camera_params) -> Tuple[torch.Tensor, torch.Tensor]:
        """get_all_poses() + get_rays_for_batch in one launch (exp map fused into ray generation)."""
        d = self.dataset
        return ops.RayGenSE3.apply(pixel_batch.image_indices, pixel_batch.pixel_coords, camera_params.initial_poses,
                                   camera_params.rotation_deltas, camera_params.translation_deltas,
                                   camera_params.learn_rotation, camera_params.learn_translation, d.H, d.W, d.focal,
                                   d.W / 2.0, d.H / 2.0)


def create_pixel_dataset(data: Any, uint8_images: bool = False) -> Tuple[PixelDataset, PixelSampler]:
    """`uint8_images=True` stores the image table as uint8 (lossless for k/255 images; raises otherwise)."""
    if uint8_images and data.images.dtype != torch.uint8:
        import copy
        data = copy.copy(data)
        data.images = ops.quantize_images(data.images)
    dataset = PixelDataset(data)
    return dataset, PixelSampler(dataset, batch_size=1024)
