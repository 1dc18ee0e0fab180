"""Clean-pose ray batches behind the reference's interface (noisy_src/data.py:160-321: RayDataset, RaySampler).

The reference precomputes three tables over all N*H*W rays (origins, directions, colours: 2.3 GB at 100 x 800^2, on top of
the images) and gathers rows from them.  Here nothing is precomputed: a batch of flat ray indices goes through the
bookkeeping kernel (image = idx // (H*W), v, u: `rn_pixel_gather`, colours gathered from the image table, fp32 or uint8)
and the ray-generation kernel (`rn_raygen_fwd`: normalise(R[img] . dir(u, v)), origin = t[img]) -- two launches, values
bit-identical to rows of the reference's tables (tests/test_gpu_parity.py::test_ray_sampler_matches_reference_tables).
`RaySampler` keeps the reference's iteration semantics: `torch.randperm` per epoch on the data's device, consecutive
slices of it, a short last batch, `sample_batch()` = `torch.randint` with replacement.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Optional

import torch

from . import ops
from .noise import NoiseConfig, add_noise_to_poses


class RayDataset:
    """All rays of a scene, generated on demand (noisy_src/data.py:160-263)."""

    def __init__(self, data: Any, batch_size: int = 1024, noise_config: Optional[NoiseConfig] = None, uint8_images: bool = False):
        self.H, self.W, self.focal = int(data.H), int(data.W), float(data.focal)
        self.noise_config = noise_config
        images = data.images
        if images.device.type != "cuda":
            raise RuntimeError("RayDataset needs CUDA-resident images (no CPU fallback)")
        if uint8_images and images.dtype != torch.uint8:
            images = ops.quantize_images(images)
        self.images = images.contiguous()
        self.original_poses = data.poses.clone()
        self.noise_info = None
        poses = data.poses
        if noise_config is not None and noise_config.has_noise:
            poses, self.noise_info = add_noise_to_poses(poses, noise_config)      # reference draw order (noise.py:194-234)
        self.poses = poses.to(self.images.device, torch.float32).contiguous()
        self.n_images = self.images.shape[0]
        self.n_rays = self.n_images * self.H * self.W
        self.device = self.images.device

    def __len__(self) -> int:
        return self.n_rays

    def gather(self, indices: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Rows `indices` of the reference's (rays_o, rays_d, colors) tables."""
        img, uv, rgb = ops.pixel_gather(indices, self.H, self.W, self.images)
        with torch.no_grad():
            ro, rd = ops.RayGen.apply(img, uv, self.poses, self.H, self.W, self.focal, self.W / 2.0, self.H / 2.0)
        return {"rays_o": ro, "rays_d": rd, "target_rgb": rgb}

    def __getitem__(self, idx) -> Dict[str, torch.Tensor]:
        if isinstance(idx, int):
            out = self.gather(torch.tensor([idx], device=self.device, dtype=torch.int64))
            return {k: v[0] for k, v in out.items()}
        return self.gather(torch.as_tensor(idx, device=self.device, dtype=torch.int64).reshape(-1))

    # the reference's tables, materialised only if somebody asks for them
    def _all(self, key: str) -> torch.Tensor:
        return self.gather(torch.arange(self.n_rays, device=self.device))[key]

    @property
    def rays_o(self) -> torch.Tensor:
        return self._all("rays_o")

    @property
    def rays_d(self) -> torch.Tensor:
        return self._all("rays_d")

    @property
    def colors(self) -> torch.Tensor:
        return self._all("target_rgb")


class RaySampler:
    """Epoch-permutation batches of rays without DataLoader overhead (noisy_src/data.py:264-321)."""

    def __init__(self, dataset: RayDataset, batch_size: int = 1024, shuffle: bool = True):
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.device = dataset.device
        self.n_rays = dataset.n_rays
        self._reset_indices()

    def _reset_indices(self) -> None:
        if self.shuffle:
            self.indices = torch.randperm(self.n_rays, device=self.device)
        else:
            self.indices = torch.arange(self.n_rays, device=self.device)
        self.current_idx = 0

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        self._reset_indices()
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        if self.current_idx >= self.n_rays:
            raise StopIteration
        end_idx = min(self.current_idx + self.batch_size, self.n_rays)
        batch_indices = self.indices[self.current_idx:end_idx]
        self.current_idx = end_idx
        return self.dataset.gather(batch_indices)

    def __len__(self) -> int:
        return (self.n_rays + self.batch_size - 1) // self.batch_size

    def sample_batch(self) -> Dict[str, torch.Tensor]:
        indices = torch.randint(0, self.n_rays, (self.batch_size,), device=self.device)
        return self.dataset.gather(indices)
