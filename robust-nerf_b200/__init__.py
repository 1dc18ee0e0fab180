"""robust-nerf_b200: B200-native (sm_100a) render-and-train hot path of ShawnnnLiu/Robust-NeRF
behind the reference's Python API (re-export list mirrors noisy_src/__init__.py:10-23 for the
symbols on the hot path).  Importable as `robust_nerf_b200` (see robust_nerf_b200.py at the repo root)."""
__version__ = "0.1.0"

from .config import NeRFConfig, ModelConfig, RenderConfig, DataConfig, TrainConfig, PoseOptConfig
from .model import NeRF, PositionalEncoding, create_nerf
from .rendering import NeRFRenderer, render_rays, raw2outputs
from .rays import (get_ray_directions, get_rays, get_rays_batch, sample_along_rays, sample_pdf,
                   sample_hierarchical)
from .pose import CameraPoseParameters
from .data_pose_opt import PixelBatch, PixelDataset, PixelSampler, create_pixel_dataset
from .data import RayDataset, RaySampler
from .train import (train_step, train_step_with_poses, render_image, render_image_with_pose, compute_psnr,
                    Trainer, render_views_sharded, evaluate)
from .synthetic import BlenderData, make_scene, lego_poses, hemisphere_poses, add_noise_to_poses
from .metrics import image_metrics, compute_ssim, compute_mse
from .noise import NoiseConfig, set_noise_seed, draw_pose_noise, compute_pose_error, compute_pose_errors_batch
from . import noise

__all__ = [
    "NeRFConfig", "ModelConfig", "RenderConfig", "DataConfig", "TrainConfig", "PoseOptConfig",
    "NeRF", "PositionalEncoding", "create_nerf", "NeRFRenderer", "render_rays", "raw2outputs",
    "get_ray_directions", "get_rays", "get_rays_batch", "sample_along_rays", "sample_pdf", "sample_hierarchical",
    "CameraPoseParameters", "PixelBatch", "PixelDataset", "PixelSampler", "create_pixel_dataset", "RayDataset", "RaySampler",
    "train_step", "train_step_with_poses", "render_image", "render_image_with_pose", "compute_psnr", "Trainer",
    "render_views_sharded", "evaluate", "BlenderData", "make_scene", "lego_poses", "hemisphere_poses", "add_noise_to_poses",
    "image_metrics", "compute_ssim", "compute_mse",
    "NoiseConfig", "set_noise_seed", "draw_pose_noise", "compute_pose_error", "compute_pose_errors_batch", "noise",
]
