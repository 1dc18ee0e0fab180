"""Volume rendering behind the reference's interface (noisy_src/rendering.py:20-323)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .config import RenderConfig
from .model import NeRF
from .rays import _z_base


def raw2outputs(rgb: torch.Tensor, sigma: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor,
                raw_noise_std: float = 0.0, white_background: bool = True, *, noise: torch.Tensor | None = None,
                early_stop_T: float = 0.0) -> Dict[str, torch.Tensor]:
    """Alpha compositing (rendering.py:20-116) -> rgb_map, depth_map, acc_map, weights."""
    if z_vals.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("no gradient w.r.t. z_vals (the reference's graph never needs one)")
    lead = z_vals.shape[:-1]
    S = z_vals.shape[-1]
    sig = sigma.reshape(*lead, S) if sigma.dim() == z_vals.dim() + 1 else sigma
    if raw_noise_std > 0.0 and noise is None:
        noise = torch.randn_like(sig) * raw_noise_std                    # rendering.py:78-80
    rgb_map, depth, acc, w = ops.Composite.apply(rgb.reshape(-1, S, 3), sig.reshape(-1, S), z_vals.reshape(-1, S),
                                                 rays_d.reshape(-1, 3), None if noise is None else noise.reshape(-1, S),
                                                 white_background, False, early_stop_T)
    return {"rgb_map": rgb_map.reshape(*lead, 3), "depth_map": depth.reshape(lead), "acc_map": acc.reshape(lead),
            "weights": w.reshape(*lead, S)}


def _run_net(model, pts_flat, viewdirs, S):
    """(B*S,3) points, (B,3) unit view dirs -> raw [B,S,4] (ours) or (rgb, sigma) (foreign module)."""
    B = viewdirs.shape[0]
    if isinstance(model, NeRF):
        return model.forward_raw(pts_flat, viewdirs, S).reshape(B, S, 4), True
    vd = viewdirs[:, None, :].expand(-1, S, -1).reshape(-1, 3)
    rgb, sigma = model(pts_flat, vd)
    return (rgb.reshape(B, S, 3), sigma.reshape(B, S)), False


def _composite(net_out, is_raw, z, rays_d, noise, white, early_stop_T):
    if is_raw:
        return ops.Composite.apply(net_out, None, z, rays_d, noise, white, True, early_stop_T)
    return ops.Composite.apply(net_out[0], net_out[1], z, rays_d, noise, white, False, early_stop_T)


def render_rays(model_coarse: NeRF, model_fine: Optional[NeRF], rays_o: torch.Tensor, rays_d: torch.Tensor,
                config: RenderConfig, is_train: bool = True, *, t_rand: torch.Tensor | None = None,
                u: torch.Tensor | None = None, early_stop_T: float = 0.0, return_extras: bool = False
                ) -> Dict[str, torch.Tensor]:
    """Coarse -> inverse-CDF resample -> fine pipeline (rendering.py:119-240)."""
    from . import rays as R
    perturb = config.perturb if is_train else False
    raw_noise_std = config.raw_noise_std if is_train else 0.0
    B = rays_o.shape[0]
    Nc = config.num_samples
    if torch.is_grad_enabled() and rays_d.requires_grad:
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)     # rendering.py:165, differentiable (pose optimisation)
    else:
        viewdirs = ops.view_dirs(rays_d)                                 # same expression, one launch
    pts_c, z_c = R.sample_along_rays(rays_o, rays_d, config.near, config.far, Nc, perturb=perturb, t_rand=t_rand)
    out_c, is_raw = _run_net(model_coarse, pts_c.reshape(-1, 3), viewdirs, Nc)
    noise_c = torch.randn(B, Nc, device=rays_o.device) * raw_noise_std if raw_noise_std > 0.0 else None
    rgb_c, depth_c, acc_c, w_c = _composite(out_c, is_raw, z_c, rays_d, noise_c, config.white_background, early_stop_T)
    results = {"rgb_coarse": rgb_c, "depth_coarse": depth_c, "acc_coarse": acc_c}
    if return_extras:
        results.update(z_coarse=z_c, weights_coarse=w_c)
    if config.use_hierarchical and model_fine is not None:
        pts_f, z_f = R.sample_hierarchical(rays_o, rays_d, z_c, w_c, config.num_samples_fine, det=not is_train, u=u)
        Nt = z_f.shape[-1]
        out_f, is_raw = _run_net(model_fine, pts_f.reshape(-1, 3), viewdirs, Nt)
        noise_f = torch.randn(B, Nt, device=rays_o.device) * raw_noise_std if raw_noise_std > 0.0 else None
        rgb_f, depth_f, acc_f, w_f = _composite(out_f, is_raw, z_f, rays_d, noise_f, config.white_background, early_stop_T)
        results.update(rgb_fine=rgb_f, depth_fine=depth_f, acc_fine=acc_f)
        if return_extras:
            results.update(z_fine=z_f, weights_fine=w_f)
    return results


class NeRFRenderer(nn.Module):
    """nn.Module owning both nets with ray chunking (rendering.py:243-323)."""

    def __init__(self, model_coarse: NeRF, model_fine: Optional[NeRF], config: RenderConfig):
        super().__init__()
        self.model_coarse = model_coarse
        self.model_fine = model_fine
        self.config = config

    def forward(self, rays_o: torch.Tensor, rays_d: torch.Tensor, chunk_size: int = 1024 * 32,
                is_train: bool = True) -> Dict[str, torch.Tensor]:
        N_rays = rays_o.shape[0]
        if N_rays <= chunk_size:
            return render_rays(self.model_coarse, self.model_fine, rays_o, rays_d, self.config, is_train=is_train)
        all_results: Dict[str, list] = {}
        for i in range(0, N_rays, chunk_size):
            res = render_rays(self.model_coarse, self.model_fine, rays_o[i:i + chunk_size], rays_d[i:i + chunk_size],
                              self.config, is_train=is_train)
            for k, v in res.items():
                all_results.setdefault(k, []).append(v)
        return {k: torch.cat(v, dim=0) for k, v in all_results.items()}
