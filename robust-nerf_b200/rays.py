"""Ray generation and sampling behind the reference's interface (noisy_src/rays.py:17-333).

Same function names, positional arguments and return values; the bodies are single sm_100a kernel
launches.  Random draws use the reference's own torch.rand calls (same shape, device and order, so
the same seed reproduces the reference's Philox stream); tests may pass the draws explicitly through
the keyword-only `t_rand=` / `u=` arguments.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("robust-nerf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def get_ray_directions(H: int, W: int, focal: float, center: Tuple[float, float] | None = None, *,
                       device=None) -> torch.Tensor:
    """(H, W, 3) camera-frame directions; no half-pixel offset, -Z forward (rays.py:17-64).
    Built on the GPU (the reference builds it on the CPU and callers `.to(device)` it)."""
    cx, cy = (W / 2.0, H / 2.0) if center is None else center
    return ops.ray_directions(H, W, focal, cx, cy, _device(device))


def get_rays(directions: torch.Tensor, c2w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rays_d = normalise(R @ dir), rays_o = t broadcast (rays.py:67-99); differentiable in c2w."""
    shp = directions.shape
    ro, rd = ops.GetRays.apply(directions.reshape(-1, 3), c2w)
    return ro.reshape(shp), rd.reshape(shp)


def get_rays_batch(H: int, W: int, focal: float, c2w_batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rays.py:102-142."""
    directions = get_ray_directions(H, W, focal, device=c2w_batch.device)
    outs = [get_rays(directions, c2w_batch[i]) for i in range(c2w_batch.shape[0])]
    return torch.stack([o for o, _ in outs], 0), torch.stack([d for _, d in outs], 0)


_Z_BASE_CACHE = {}


def _z_base(near, far, num_samples, lindisp, device):
    """rays.py:185-192, evaluated with the same torch ops on the same device; constant per (near, far, N, lindisp, device),
    so it is built once (four launches) instead of on every call."""
    key = (float(near), float(far), int(num_samples), bool(lindisp), str(device))
    z = _Z_BASE_CACHE.get(key)
    if z is None:
        t_vals = torch.linspace(0.0, 1.0, num_samples, device=device)
        if lindisp:
            z = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
        else:
            z = near * (1.0 - t_vals) + far * t_vals
        if not torch.cuda.is_current_stream_capturing():
            _Z_BASE_CACHE[key] = z
    return z


def sample_along_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, near: float, far: float, num_samples: int,
                      perturb: bool = True, lindisp: bool = False, *, t_rand: torch.Tensor | None = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stratified sampling (rays.py:145-210) -> pts (..., N, 3), z_vals (..., N)."""
    device = rays_o.device
    batch_shape = rays_o.shape[:-1]
    zb = _z_base(near, far, num_samples, lindisp, device)
    if perturb and t_rand is None:
        t_rand = torch.rand(*batch_shape, num_samples, device=device)      # rays.py:204
    o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    tr = None if not perturb else t_rand.reshape(-1, num_samples)
    need_grad = torch.is_grad_enabled() and (rays_o.requires_grad or rays_d.requires_grad)
    z, pts = ops.stratified(o.detach(), d.detach(), zb, tr, want_pts=not need_grad)
    if need_grad:
        pts = ops.Points.apply(o, d, z)
    return pts.reshape(*batch_shape, num_samples, 3), z.reshape(*batch_shape, num_samples)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, num_samples: int, det: bool = False, *,
               u: torch.Tensor | None = None) -> torch.Tensor:
    """Inverse-CDF sampling (rays.py:213-279)."""
    device = weights.device
    lead = weights.shape[:-1]
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, num_samples, device=device)      # shared row, rays.py:252
        else:
            u = torch.rand(*lead, num_samples, device=device)             # rays.py:255
    if u.dim() > 1:
        u = u.reshape(-1, num_samples)
    out = ops.sample_pdf(bins.detach().reshape(-1, bins.shape[-1]), weights.detach().reshape(-1, weights.shape[-1]), u)
    return out.reshape(*lead, num_samples)


def sample_hierarchical(rays_o: torch.Tensor, rays_d: torch.Tensor, z_vals: torch.Tensor, weights: torch.Tensor,
                        num_samples_fine: int, det: bool = False, *, u: torch.Tensor | None = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Hierarchical resampling + sorted merge (rays.py:282-333) -> pts (..., Nc+Nf, 3), z (..., Nc+Nf)."""
    device = weights.device
    lead = z_vals.shape[:-1]
    Nc = z_vals.shape[-1]
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, num_samples_fine, device=device)
        else:
            u = torch.rand(*lead, num_samples_fine, device=device)
    if u.dim() > 1:
        u = u.reshape(-1, num_samples_fine)
    o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    need_grad = torch.is_grad_enabled() and (rays_o.requires_grad or rays_d.requires_grad)
    z_all, pts, _ = ops.sample_hierarchical(o.detach(), d.detach(), z_vals.detach().reshape(-1, Nc),
                                            weights.detach().reshape(-1, Nc), u, want_pts=not need_grad)
    if need_grad:
        pts = ops.Points.apply(o, d, z_all)
    Nt = Nc + num_samples_fine
    return pts.reshape(*lead, Nt, 3), z_all.reshape(*lead, Nt)
