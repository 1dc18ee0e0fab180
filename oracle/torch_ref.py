"""Plain-PyTorch fp32 restatement of the Robust-NeRF render-and-train hot path.

TEST INFRASTRUCTURE ONLY (same rule as `oracle/nerf_oracle.py`): only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s baseline legs (`cpu_baseline`, `--impl reference`,
`--impl torch-eager`) may import this module; the product package never does.

Why a second oracle: the numpy oracle finishes a few hundred rays in seconds; this one runs on
whatever device its tensors live on, so

* on the B200 it is the fp32 comparison point at the benchmark's own size (4096 rays, 1 M
  points, autograd gradients) and the like-for-like "eager PyTorch on the same GPU" baseline
  (last sentence of the CPU-baseline row);
* on the host cores it is the CPU baseline: the same aten operators the reference calls
  (`addmm`, `cumprod`, `searchsorted`, `sort`, autograd), threaded the same way.

It is written from the algorithm description (SURVEY 8a + quirk ledger), not from the reference's
source text, and is PINNED in `tests/test_torch_ref_golden.py` against the golden vectors the
unmodified reference produced (`tests/golden/*.npz`): forward values bit-for-bit or to 1e-6,
autograd gradients to 1e-5.  Each function cites the reference lines it restates (paths relative
to the reference repository).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]

POS_FREQS, DIR_FREQS, HIDDEN, N_LAYERS, SKIP_AFTER = 10, 4, 256, 8, 4


# ------------------------------------------------------------------------------------------------
# model.py
# ------------------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, num_freqs: int) -> torch.Tensor:
    """noisy_src/model.py:58-80: [x, sin(2^k x), cos(2^k x)] for k = 0..L-1, no pi (quirk 1)."""
    feats = [x]
    for k in range(num_freqs):
        f = float(2 ** k)
        feats += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(feats, dim=-1)


def param_names() -> Sequence[str]:
    names = []
    for i in range(N_LAYERS):
        names += [f"pts_linears.{i}.weight", f"pts_linears.{i}.bias"]
    for n in ("sigma_linear", "feature_linear", "dir_linear", "rgb_linear"):
        names += [f"{n}.weight", f"{n}.bias"]
    return names


def to_params(weights, device, requires_grad: bool = True) -> Params:
    """numpy / tensor weight dict (state_dict names) -> fp32 leaf tensors on `device`."""
    out = {}
    for k in param_names():
        t = torch.as_tensor(weights[k]).detach().to(device=device, dtype=torch.float32).clone()
        out[k] = t.requires_grad_(requires_grad)
    return out


def nerf_forward(p: Params, x: torch.Tensor, d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """noisy_src/model.py:145-196.  x, d (M,3) -> rgb (M,3), sigma (M,1).
    Skip concat puts x_enc FIRST and happens after layer 4's ReLU (quirk 2); the feature layer has no
    activation (quirk 3)."""
    lin = torch.nn.functional.linear
    x_enc = positional_encoding(x, POS_FREQS)
    h = x_enc
    for i in range(N_LAYERS):
        h = torch.relu(lin(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i == SKIP_AFTER:
            h = torch.cat([x_enc, h], dim=-1)
    sigma = torch.relu(lin(h, p["sigma_linear.weight"], p["sigma_linear.bias"]))
    feat = lin(h, p["feature_linear.weight"], p["feature_linear.bias"])
    hc = torch.relu(lin(torch.cat([feat, positional_encoding(d, DIR_FREQS)], dim=-1),
                        p["dir_linear.weight"], p["dir_linear.bias"]))
    rgb = torch.sigmoid(lin(hc, p["rgb_linear.weight"], p["rgb_linear.bias"]))
    return rgb, sigma


# ------------------------------------------------------------------------------------------------
# rays.py
# ------------------------------------------------------------------------------------------------
def get_ray_directions(H: int, W: int, focal: float, center=None, device="cpu") -> torch.Tensor:
    """noisy_src/rays.py:17-64: no half-pixel offset, principal point (W/2, H/2), -Z forward (quirk 5).
    Built on the CPU like the reference, then moved."""
    cx, cy = (W / 2.0, H / 2.0) if center is None else center
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    dirs = torch.stack([(i - cx) / focal, -(j - cy) / focal, -torch.ones_like(i)], dim=-1)
    return dirs.to(device)


def get_rays(directions: torch.Tensor, c2w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """noisy_src/rays.py:67-99: rotate, normalise; origin = translation column."""
    rays_d = torch.sum(directions[..., None, :] * c2w[:3, :3], dim=-1)
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d


def sample_along_rays(rays_o, rays_d, near, far, num_samples, perturb=True, lindisp=False, t_rand=None):
    """noisy_src/rays.py:145-210."""
    t = torch.linspace(0.0, 1.0, num_samples, device=rays_o.device)
    z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t) if lindisp else near * (1.0 - t) + far * t
    z = z.expand(*rays_o.shape[:-1], num_samples)
    if perturb:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], dim=-1)
        lower = torch.cat([z[..., :1], mids], dim=-1)
        if t_rand is None:
            t_rand = torch.rand(*rays_o.shape[:-1], num_samples, device=rays_o.device)
        z = lower + (upper - lower) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return pts, z


def sample_pdf(bins, weights, num_samples, det=False, u=None, return_inds=False):
    """noisy_src/rays.py:213-279: +1e-5, cdf with a leading 0, searchsorted(right=True), clamped gathers,
    denom < 1e-5 -> 1 (quirk 9)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, dim=-1)], dim=-1)
    if det:
        u = torch.linspace(0.0, 1.0, num_samples, device=bins.device).expand(*cdf.shape[:-1], num_samples)
    elif u is None:
        u = torch.rand(*cdf.shape[:-1], num_samples, device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    samples = bin_b + (u - cdf_b) / denom * (bin_a - bin_b)
    return (samples, inds, cdf) if return_inds else samples


def sample_hierarchical(rays_o, rays_d, z_vals, weights, num_samples_fine, det=False, u=None):
    """noisy_src/rays.py:282-333: mid-point bins, interior weights, detached samples, sorted merge (quirk 10)."""
    mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    z_new = sample_pdf(mids, weights[..., 1:-1], num_samples_fine, det=det, u=u).detach()
    z_all, _ = torch.sort(torch.cat([z_vals, z_new], dim=-1), dim=-1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
    return pts, z_all


# ------------------------------------------------------------------------------------------------
# rendering.py
# ------------------------------------------------------------------------------------------------
def raw2outputs(rgb, sigma, z_vals, rays_d, noise=None, white_background=True):
    """noisy_src/rendering.py:20-116: last interval 1e10, transmittance over (1 - alpha + 1e-10) (quirk 8)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], dim=-1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    s = sigma[..., 0]
    if noise is not None:
        s = s + noise
    alpha = 1.0 - torch.exp(-torch.relu(s) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], dim=-1), dim=-1)[..., :-1]
    w = alpha * trans
    rgb_map = torch.sum(w[..., None] * rgb, dim=-2)
    depth = torch.sum(w * z_vals, dim=-1)
    acc = torch.sum(w, dim=-1)
    if white_background:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    return {"rgb_map": rgb_map, "depth_map": depth, "acc_map": acc, "weights": w}


def render_rays(pc: Params, pf: Optional[Params], rays_o, rays_d, near=2.0, far=6.0, num_samples=64,
                num_samples_fine=128, is_train=True, perturb=True, white_background=True, t_rand=None, u=None):
    """noisy_src/rendering.py:119-240 (raw_noise_std = 0, the reference's only configuration)."""
    B = rays_o.shape[0]
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    pts, z = sample_along_rays(rays_o, rays_d, near, far, num_samples, perturb=perturb and is_train, t_rand=t_rand)
    vd = viewdirs[:, None, :].expand(-1, num_samples, -1).reshape(-1, 3)
    rgb, sigma = nerf_forward(pc, pts.reshape(-1, 3), vd)
    oc = raw2outputs(rgb.reshape(B, num_samples, 3), sigma.reshape(B, num_samples, 1), z, rays_d, None, white_background)
    res = {"rgb_coarse": oc["rgb_map"], "depth_coarse": oc["depth_map"], "acc_coarse": oc["acc_map"],
           "z_coarse": z, "weights_coarse": oc["weights"]}
    if pf is not None:
        pts_f, z_f = sample_hierarchical(rays_o, rays_d, z, oc["weights"], num_samples_fine, det=not is_train, u=u)
        nt = z_f.shape[-1]
        vd = viewdirs[:, None, :].expand(-1, nt, -1).reshape(-1, 3)
        rgb, sigma = nerf_forward(pf, pts_f.reshape(-1, 3), vd)
        of = raw2outputs(rgb.reshape(B, nt, 3), sigma.reshape(B, nt, 1), z_f, rays_d, None, white_background)
        res.update(rgb_fine=of["rgb_map"], depth_fine=of["depth_map"], acc_fine=of["acc_map"], z_fine=z_f,
                   weights_fine=of["weights"])
    return res


def render_loss(res, target):
    """noisy_src/train.py:88-99: mse(coarse) + mse(fine)."""
    loss = torch.mean((res["rgb_coarse"] - target) ** 2)
    if "rgb_fine" in res:
        loss = loss + torch.mean((res["rgb_fine"] - target) ** 2)
    return loss


# ------------------------------------------------------------------------------------------------
# train_pose_opt.py :: CameraPoseParameters, data_pose_opt.py :: get_rays_from_pixels
# ------------------------------------------------------------------------------------------------
def axis_angle_to_rotation_matrix(w: torch.Tensor) -> torch.Tensor:
    """noisy_src/train_pose_opt.py:122-184 (Rodrigues; |w| < 1e-6 selects the constant identity, so the
    gradient of that branch is exactly zero: quirk 11)."""
    theta = torch.norm(w, dim=-1, keepdim=True)
    small = theta < 1e-6
    theta_s = torch.where(small, torch.ones_like(theta), theta)
    k = w / theta_s
    zero = torch.zeros_like(k[..., 0])
    K = torch.stack([zero, -k[..., 2], k[..., 1], k[..., 2], zero, -k[..., 0], -k[..., 1], k[..., 0], zero],
                    dim=-1).reshape(*w.shape[:-1], 3, 3)
    eye = torch.eye(3, device=w.device, dtype=w.dtype).expand_as(K)
    s, c = torch.sin(theta_s)[..., None], torch.cos(theta_s)[..., None]
    R = eye + s * K + (1.0 - c) * (K @ K)
    return torch.where(small[..., None], eye, R)


def get_poses(initial_poses, rot_deltas, trans_deltas) -> torch.Tensor:
    """noisy_src/train_pose_opt.py:186-230: R = R_delta @ R_init, t = t_init + dt, last row (0,0,0,1)."""
    R = axis_angle_to_rotation_matrix(rot_deltas) @ initial_poses[:, :3, :3]
    t = initial_poses[:, :3, 3] + trans_deltas
    top = torch.cat([R, t[..., None]], dim=-1)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], device=top.device).expand(top.shape[0], 1, 4)
    return torch.cat([top, bottom], dim=1)


def rays_from_pixels(image_indices, pixel_coords, poses, directions):
    """Net effect of noisy_src/data_pose_opt.py:83-148,200-223 (quirks 14, 15): per-pixel direction looked up at
    (v, u) = pixel_coords[:, 1], [:, 0] cast to long, rotated by the pixel's own pose, normalised."""
    d = directions[pixel_coords[:, 1].long(), pixel_coords[:, 0].long()]
    P = poses[image_indices]
    rays_d = torch.sum(d[:, None, :] * P[:, :3, :3], dim=-1)
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    return P[:, :3, 3], rays_d


# ------------------------------------------------------------------------------------------------
# training steps (train.py:68-119, train_pose_opt.py:290-411)
# ------------------------------------------------------------------------------------------------
class RefTrainer:
    """Eager reference-semantics training on any device.  Clean mode: joint clip at 1.0 over both nets, one Adam.
    Pose mode: per-net clip 1.0, joint pose clip 0.1, second Adam at pose_lr, optional L2 pose regulariser (quirk 13).
    LR schedule: lr * 0.1^(step / 250000) (train.py:405-411)."""

    def __init__(self, pc: Params, pf: Optional[Params], lr=5e-4, lr_decay_steps=250000.0, initial_poses=None,
                 rot=None, trans=None, pose_lr=1e-4, rot_reg=0.0, trans_reg=0.0, **render_kw):
        self.pc, self.pf, self.kw = pc, pf, render_kw
        self.net_params = list(pc.values()) + (list(pf.values()) if pf is not None else [])
        self.opt = torch.optim.Adam(self.net_params, lr=lr)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda s: 0.1 ** (s / lr_decay_steps))
        self.initial_poses, self.rot, self.trans = initial_poses, rot, trans
        self.rot_reg, self.trans_reg = rot_reg, trans_reg
        self.opt_pose = self.sched_pose = None
        if rot is not None:
            self.opt_pose = torch.optim.Adam([rot, trans], lr=pose_lr)
            self.sched_pose = torch.optim.lr_scheduler.LambdaLR(self.opt_pose, lambda s: 0.1 ** (s / lr_decay_steps))

    def grads_rays(self, rays_o, rays_d, target, t_rand=None, u=None):
        for p in self.net_params:
            p.grad = None
        res = render_rays(self.pc, self.pf, rays_o, rays_d, is_train=True, t_rand=t_rand, u=u, **self.kw)
        loss = render_loss(res, target)
        loss.backward()
        return loss.detach(), res

    def step_rays(self, rays_o, rays_d, target, t_rand=None, u=None):
        loss, _ = self.grads_rays(rays_o, rays_d, target, t_rand, u)
        torch.nn.utils.clip_grad_norm_(self.net_params, max_norm=1.0)
        self.opt.step()
        self.sched.step()
        return loss

    def step_pixels(self, image_indices, pixel_coords, target, directions, optimize_poses=True, t_rand=None, u=None):
        for p in self.net_params + [self.rot, self.trans]:
            p.grad = None
        poses = get_poses(self.initial_poses, self.rot, self.trans)
        rays_o, rays_d = rays_from_pixels(image_indices, pixel_coords, poses, directions)
        res = render_rays(self.pc, self.pf, rays_o, rays_d, is_train=True, t_rand=t_rand, u=u, **self.kw)
        loss = render_loss(res, target)
        if optimize_poses:
            if self.rot_reg > 0:
                loss = loss + self.rot_reg * torch.mean(self.rot ** 2)
            if self.trans_reg > 0:
                loss = loss + self.trans_reg * torch.mean(self.trans ** 2)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.pc.values()), max_norm=1.0)
        if self.pf is not None:
            torch.nn.utils.clip_grad_norm_(list(self.pf.values()), max_norm=1.0)
        self.opt.step()
        self.sched.step()
        if optimize_poses:
            torch.nn.utils.clip_grad_norm_([self.rot, self.trans], max_norm=0.1)
            self.opt_pose.step()
            self.sched_pose.step()
        return loss.detach()


@torch.no_grad()
def render_image(pc: Params, pf: Optional[Params], pose, H, W, focal, chunk=4096, **kw):
    """noisy_src/train.py:122-160 / rendering.py:287-323: full image in ray chunks, eval mode."""
    dirs = get_ray_directions(H, W, focal, device=pose.device)
    ro, rd = get_rays(dirs, pose)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    outs = []
    for a in range(0, ro.shape[0], chunk):
        r = render_rays(pc, pf, ro[a:a + chunk], rd[a:a + chunk], is_train=False, **kw)
        outs.append(r["rgb_fine"] if "rgb_fine" in r else r["rgb_coarse"])
    return torch.cat(outs, 0).reshape(H, W, 3)


def lego_focal(W: int = 800) -> float:
    return 0.5 * W / math.tan(0.5 * 0.6911112070083618)
