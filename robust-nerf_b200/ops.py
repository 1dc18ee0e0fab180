"""torch.autograd.Function wrappers over the C ABI (one per kernel family).

Forward and backward both run hand-written sm_100a kernels; nothing here computes on the CPU or
through aten math ops except trivial O(B) glue (e.g. normalising B view directions).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib as L
from ._lib import call, ptr, require_cuda, stream_ptr


def _f32(t, name):
    return require_cuda(t, name, torch.float32)


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


# --------------------------------------------------------------------------------------------
# SE(3) pose parameters (train_pose_opt.py:122-226)
# --------------------------------------------------------------------------------------------
class SE3Poses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, initial_poses, rot, trans, indices, learn_r, learn_t):
        P0, rot, trans = _f32(initial_poses, "initial_poses"), _f32(rot, "rotation_deltas"), _f32(trans, "translation_deltas")
        n_total = P0.shape[0]
        idx = None if indices is None else require_cuda(indices, "indices", torch.int64)
        n = n_total if idx is None else idx.numel()
        out = _empty((n, 4, 4), P0)
        call("rn_se3_poses_fwd", ptr(P0), ptr(rot), ptr(trans), ptr(idx), n, n_total, int(learn_r), int(learn_t),
             ptr(out), stream_ptr())
        ctx.save_for_backward(P0, rot, idx if idx is not None else torch.empty(0, device=P0.device, dtype=torch.int64))
        ctx.meta = (n, n_total, idx is not None, bool(learn_r), bool(learn_t))
        return out

    @staticmethod
    def backward(ctx, g):
        P0, rot, idx = ctx.saved_tensors
        n, n_total, has_idx, learn_r, learn_t = ctx.meta
        g = _f32(g, "grad_poses")
        d_rot = torch.zeros_like(rot) if (learn_r and ctx.needs_input_grad[1]) else None
        d_trans = torch.zeros((n_total, 3), device=P0.device) if (learn_t and ctx.needs_input_grad[2]) else None
        if d_rot is not None or d_trans is not None:
            call("rn_se3_poses_bwd", ptr(P0), ptr(rot), ptr(idx) if has_idx else None, n, n_total, ptr(g), ptr(d_rot),
                 ptr(d_trans), stream_ptr())
        return None, d_rot, d_trans, None, None, None


# --------------------------------------------------------------------------------------------
# ray generation (rays.py:17-99, data_pose_opt.py:83-148)
# --------------------------------------------------------------------------------------------
def ray_directions(H, W, focal, cx, cy, device):
    out = torch.empty((H, W, 3), device=device, dtype=torch.float32)
    call("rn_ray_directions", int(H), int(W), float(focal), float(cx), float(cy), ptr(out), stream_ptr())
    return out


class GetRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, directions, c2w):
        d = _f32(directions, "directions")
        c = _f32(c2w, "c2w")
        n = d.numel() // 3
        ro, rd = _empty(d.shape, d), _empty(d.shape, d)
        call("rn_get_rays", ptr(d), ptr(c), n, ptr(ro), ptr(rd), stream_ptr())
        ctx.save_for_backward(d, c)
        return ro, rd

    @staticmethod
    def backward(ctx, go, gd):
        d, c = ctx.saved_tensors
        n = d.numel() // 3
        go = _f32(go, "g_rays_o") if go is not None else torch.zeros_like(d)
        gd = _f32(gd, "g_rays_d") if gd is not None else torch.zeros_like(d)
        g_c = _empty((4, 4), d)
        g_d = _empty(d.shape, d) if ctx.needs_input_grad[0] else None
        call("rn_get_rays_bwd", ptr(d), ptr(c), n, ptr(go), ptr(gd), ptr(g_c), ptr(g_d), stream_ptr())
        return g_d, g_c


class RayGen(torch.autograd.Function):
    """pixel batch + poses[N,4,4] -> rays (gather fused)."""

    @staticmethod
    def forward(ctx, image_idx, pixel_uv, poses, H, W, focal, cx, cy):
        img = require_cuda(image_idx, "image_indices", torch.int64)
        uv = _f32(pixel_uv, "pixel_coords")
        P = _f32(poses, "poses")
        B = img.numel()
        ro, rd = _empty((B, 3), P), _empty((B, 3), P)
        call("rn_raygen_fwd", ptr(img), ptr(uv), B, ptr(P), P.shape[0], H, W, focal, cx, cy, ptr(ro), ptr(rd), stream_ptr())
        ctx.save_for_backward(img, uv, P)
        ctx.meta = (H, W, focal, cx, cy)
        return ro, rd

    @staticmethod
    def backward(ctx, go, gd):
        img, uv, P = ctx.saved_tensors
        H, W, focal, cx, cy = ctx.meta
        B = img.numel()
        go = _f32(go, "g_rays_o") if go is not None else torch.zeros((B, 3), device=P.device)
        gd = _f32(gd, "g_rays_d") if gd is not None else torch.zeros((B, 3), device=P.device)
        gP = _empty(P.shape, P)
        call("rn_raygen_bwd", ptr(img), ptr(uv), B, ptr(P), P.shape[0], H, W, focal, cx, cy, ptr(go), ptr(gd), ptr(gP),
             stream_ptr())
        return None, None, gP, None, None, None, None, None


class RayGenSE3(torch.autograd.Function):
    """Fused exp-map pose update + ray generation, backward straight into (rot, trans) deltas."""

    @staticmethod
    def forward(ctx, image_idx, pixel_uv, initial_poses, rot, trans, learn_r, learn_t, H, W, focal, cx, cy):
        img = require_cuda(image_idx, "image_indices", torch.int64)
        uv = _f32(pixel_uv, "pixel_coords")
        P0, rot, trans = _f32(initial_poses, "initial_poses"), _f32(rot, "rotation_deltas"), _f32(trans, "translation_deltas")
        B = img.numel()
        ro, rd = _empty((B, 3), P0), _empty((B, 3), P0)
        call("rn_raygen_se3_fwd", ptr(img), ptr(uv), B, ptr(P0), ptr(rot), ptr(trans), P0.shape[0], int(learn_r),
             int(learn_t), H, W, focal, cx, cy, ptr(ro), ptr(rd), stream_ptr())
        ctx.save_for_backward(img, uv, P0, rot)
        ctx.meta = (bool(learn_r), bool(learn_t), H, W, focal, cx, cy)
        return ro, rd

    @staticmethod
    def backward(ctx, go, gd):
        img, uv, P0, rot = ctx.saved_tensors
        learn_r, learn_t, H, W, focal, cx, cy = ctx.meta
        B = img.numel()
        go = _f32(go, "g_rays_o") if go is not None else torch.zeros((B, 3), device=P0.device)
        gd = _f32(gd, "g_rays_d") if gd is not None else torch.zeros((B, 3), device=P0.device)
        d_rot, d_trans = _empty(rot.shape, rot), _empty(rot.shape, rot)
        call("rn_raygen_se3_bwd", ptr(img), ptr(uv), B, ptr(P0), ptr(rot), P0.shape[0], int(learn_r), H, W, focal, cx, cy,
             ptr(go), ptr(gd), ptr(d_rot), ptr(d_trans), stream_ptr())
        return (None, None, None, d_rot if (learn_r and ctx.needs_input_grad[3]) else None,
                d_trans if (learn_t and ctx.needs_input_grad[4]) else None, None, None, None, None, None, None, None)


def pixel_gather(flat_idx, H, W, images=None):
    flat = require_cuda(flat_idx, "flat_idx", torch.int64)
    B = flat.numel()
    img = torch.empty(B, device=flat.device, dtype=torch.int64)
    uv = torch.empty((B, 2), device=flat.device, dtype=torch.float32)
    rgb = None
    if images is not None and images.dtype == torch.uint8:
        # uint8 image table: rgb = k / 255 in fp32 inside the gather (exactly the reference's conversion, data.py:134-136)
        images = require_cuda(images, "images", torch.uint8)
        rgb = torch.empty((B, 3), device=flat.device, dtype=torch.float32)
        call("rn_pixel_gather_u8", ptr(flat), B, int(H), int(W), ptr(images), ptr(img), ptr(uv), ptr(rgb), stream_ptr())
        return img, uv, rgb
    if images is not None:
        images = _f32(images, "images")
        rgb = torch.empty((B, 3), device=flat.device, dtype=torch.float32)
    call("rn_pixel_gather", ptr(flat), B, int(H), int(W), ptr(images), ptr(img), ptr(uv), ptr(rgb), stream_ptr())
    return img, uv, rgb


def quantize_images(images: torch.Tensor) -> torch.Tensor:
    """fp32 images in [0,1] -> uint8, refusing anything that is not exactly k/255 (lossless by construction for data
    loaded the reference's way, data.py:118-136)."""
    q = torch.round(images * 255.0).clamp_(0, 255).to(torch.uint8)
    if not torch.equal(dequantize_images(q), images.to(torch.float32)):
        raise ValueError("images are not exact multiples of 1/255: uint8 storage would not be lossless")
    return q


def dequantize_images(q: torch.Tensor) -> torch.Tensor:
    """uint8 -> fp32 k / 255 with a correctly rounded DIVISION, as numpy does in the reference (data.py:134-136) and as
    the gather kernel does.  (torch's CUDA `tensor / python_scalar` multiplies by the reciprocal, which differs in the
    last bit for some k; a tensor divisor takes the true-division path.)"""
    return q.to(torch.float32) / torch.full((), 255.0, device=q.device, dtype=torch.float32)


# --------------------------------------------------------------------------------------------
# sampling (rays.py:145-333)
# --------------------------------------------------------------------------------------------
class Points(torch.autograd.Function):
    """pts = o + d * z  (z carries no gradient in the reference's graph: rays.py:325)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z):
        o, d, z = _f32(rays_o, "rays_o"), _f32(rays_d, "rays_d"), _f32(z, "z_vals")
        B, S = z.shape
        pts = _empty((B, S, 3), z)
        call("rn_points_fwd", ptr(o), ptr(d), ptr(z), B, S, ptr(pts), stream_ptr())
        ctx.save_for_backward(z)
        return pts

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        B, S = z.shape
        g = _f32(g, "g_pts")
        go, gd = _empty((B, 3), z), _empty((B, 3), z)
        call("rn_points_bwd", ptr(g), ptr(z), B, S, ptr(go), ptr(gd), stream_ptr())
        return go, gd, None


def stratified(rays_o, rays_d, z_base, t_rand, want_pts=True):
    """z (and pts) of rays.py:197-208; no autograd (use Points for the differentiable pts)."""
    o, d = _f32(rays_o, "rays_o"), _f32(rays_d, "rays_d")
    zb = _f32(z_base, "z_base")
    B, Nc = o.shape[0], zb.numel()
    tr = None if t_rand is None else _f32(t_rand, "t_rand")
    z = _empty((B, Nc), o)
    pts = _empty((B, Nc, 3), o) if want_pts else None
    call("rn_stratified_fwd", ptr(o), ptr(d), B, ptr(zb), Nc, ptr(tr), ptr(z), ptr(pts), stream_ptr())
    return z, pts


def sample_pdf(bins, weights, u, return_inds=False):
    bins, weights, u = _f32(bins, "bins"), _f32(weights, "weights"), _f32(u, "u")
    B, nb = bins.shape
    Nf = u.shape[-1]
    stride = 0 if u.dim() == 1 else Nf
    out = _empty((B, Nf), bins)
    inds = torch.empty((B, Nf), device=bins.device, dtype=torch.int64) if return_inds else None
    call("rn_sample_pdf_fwd", ptr(bins), ptr(weights), B, nb, ptr(u), stride, Nf, ptr(out), ptr(inds), stream_ptr())
    return (out, inds) if return_inds else out


def sample_hierarchical(rays_o, rays_d, z_coarse, weights, u, want_pts=True, return_inds=False):
    o, d = _f32(rays_o, "rays_o"), _f32(rays_d, "rays_d")
    zc, w, u = _f32(z_coarse, "z_vals"), _f32(weights, "weights"), _f32(u, "u")
    B, Nc = zc.shape
    Nf = u.shape[-1]
    stride = 0 if u.dim() == 1 else Nf
    z_all = _empty((B, Nc + Nf), zc)
    pts = _empty((B, Nc + Nf, 3), zc) if want_pts else None
    inds = torch.empty((B, Nf), device=zc.device, dtype=torch.int64) if return_inds else None
    call("rn_sample_hierarchical_fwd", ptr(o), ptr(d), ptr(zc), ptr(w), B, Nc, ptr(u), stride, Nf, ptr(z_all), ptr(pts),
         ptr(inds), stream_ptr())
    return z_all, pts, inds


# --------------------------------------------------------------------------------------------
# compositing (rendering.py:20-116)
# --------------------------------------------------------------------------------------------
class Composite(torch.autograd.Function):
    """raw_mode=False: inputs (rgb[B,S,3], sigma[B,S]); raw_mode=True: a = raw4[B,S,4], b unused."""

    @staticmethod
    def forward(ctx, a, b, z, rays_d, noise, white, raw_mode, t_min):
        z, rd = _f32(z, "z_vals"), _f32(rays_d, "rays_d")
        B, S = z.shape
        if raw_mode:
            raw4, rgb, sigma = _f32(a, "raw"), None, None
        else:
            raw4, rgb, sigma = None, _f32(a, "rgb"), _f32(b, "sigma")
        noise = None if noise is None else _f32(noise, "noise")
        rgb_map, depth, acc, w = _empty((B, 3), z), _empty((B,), z), _empty((B,), z), _empty((B, S), z)
        call("rn_composite_fwd", ptr(rgb), ptr(sigma), ptr(raw4), ptr(z), ptr(rd), ptr(noise), B, S, int(white),
             float(t_min), ptr(rgb_map), ptr(depth), ptr(acc), ptr(w), stream_ptr())
        ctx.save_for_backward(*(t for t in (raw4, rgb, sigma, z, rd, noise) if t is not None))
        ctx.meta = (raw_mode, noise is not None, bool(white))
        # unused outputs (depth, acc, weights in training) arrive as None in backward instead of zero tensors: no fill
        # launches, and the backward kernel takes its lean variant
        ctx.set_materialize_grads(False)
        return rgb_map, depth, acc, w

    @staticmethod
    def backward(ctx, g_map, g_depth, g_acc, g_w):
        raw_mode, has_noise, white = ctx.meta
        sv = list(ctx.saved_tensors)
        raw4 = sv.pop(0) if raw_mode else None
        rgb = None if raw_mode else sv.pop(0)
        sigma = None if raw_mode else sv.pop(0)
        z, rd = sv.pop(0), sv.pop(0)
        noise = sv.pop(0) if has_noise else None
        B, S = z.shape
        g_map = _f32(g_map, "g_rgb_map") if g_map is not None else torch.zeros((B, 3), device=z.device)
        g_depth = None if g_depth is None else _f32(g_depth, "g_depth")
        g_acc = None if g_acc is None else _f32(g_acc, "g_acc")
        g_w = None if g_w is None else _f32(g_w, "g_weights")
        d_raw = _empty((B, S, 4), z) if raw_mode else None
        d_rgb = None if raw_mode else _empty((B, S, 3), z)
        d_sigma = None if raw_mode else _empty((B, S), z)
        d_rd = _empty((B, 3), z) if ctx.needs_input_grad[3] else None
        call("rn_composite_bwd", ptr(rgb), ptr(sigma), ptr(raw4), ptr(z), ptr(rd), ptr(noise), B, S, int(white), ptr(g_map),
             ptr(g_depth), ptr(g_acc), ptr(g_w), ptr(d_rgb), ptr(d_sigma), ptr(d_raw), ptr(d_rd), stream_ptr())
        if raw_mode:
            return d_raw, None, None, d_rd, None, None, None, None
        return d_rgb, d_sigma, None, d_rd, None, None, None, None


def mse_loss_and_grad(rgb_map, target, loss_out, scale=1.0, want_grad=True):
    rgb_map, target = _f32(rgb_map, "rgb_map"), _f32(target, "target")
    B = rgb_map.shape[0]
    g = _empty((B, 3), rgb_map) if want_grad else None
    call("rn_mse_loss_fwd_bwd", ptr(rgb_map), ptr(target), B, float(scale), ptr(loss_out), ptr(g), stream_ptr())
    return g


def mse2_loss_and_grads(rgb_coarse, rgb_fine, target):
    """loss[0] = mse(coarse) + mse(fine), loss[1] = coarse, loss[2] = fine and both gradients w.r.t. the rendered colours,
    one launch (train.py:88-99).  rgb_fine may be None."""
    rc, tg = _f32(rgb_coarse.detach(), "rgb_coarse"), _f32(target, "target")
    rf = None if rgb_fine is None else _f32(rgb_fine.detach(), "rgb_fine")
    B = rc.shape[0]
    loss = _empty((3,), rc)
    g_c = _empty((B, 3), rc)
    g_f = None if rf is None else _empty((B, 3), rc)
    call("rn_mse2_loss_fwd_bwd", ptr(rc), ptr(rf), ptr(tg), B, ptr(loss), ptr(g_c), ptr(g_f), stream_ptr())
    return loss, g_c, g_f


# --------------------------------------------------------------------------------------------
# NeRF MLP (model.py:145-196)
# --------------------------------------------------------------------------------------------
class PackedWeights:
    """bf16 padded weight cache of one NeRF; refreshed when any parameter changed
    (data_ptr or in-place version counter, e.g. after an optimiser step)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.key = None
        # Optional flat fp32 gradient sink (RN_NUM_PARAMS floats).  While set (only inside a Trainer step, see
        # Trainer._sinks), the backward kernels write this network's parameter gradients straight into it and autograd
        # receives None for the parameters -- no 48 accumulate launches.  The first backward through the net after
        # the sink was armed overwrites, any further one (chunked rendering, gradient accumulation) adds.
        self.grad_sink: Optional[torch.Tensor] = None
        self.sink_dirty = False
        self.on_grads_ready = None             # Trainer hook: called after this net's gradients landed in the sink

    def get(self, params: Sequence[torch.Tensor]) -> torch.Tensor:
        key = tuple((p.data_ptr(), p._version) for p in params)
        dev = params[0].device
        if self.buf is None or self.buf.device != dev:
            self.buf = torch.empty(L.lib().rn_mlp_packed_weight_bytes(), device=dev, dtype=torch.uint8)
            self.key = None
        if key != self.key:
            arr = (ctypes.c_void_p * L.NUM_PARAM_TENSORS)(*[p.data_ptr() for p in params])
            call("rn_mlp_pack_weights", arr, ptr(self.buf), stream_ptr())
            self.key = key
        return self.buf


_PARAM_SHAPES = ([(256, 63), (256,)] + [(256, 256), (256,)] * 4 + [(256, 319), (256,)] + [(256, 256), (256,)] * 2
                 + [(1, 256), (1,), (256, 256), (256,), (128, 283), (128,), (3, 128), (3,)])


def split_flat_grads(flat: torch.Tensor):
    out, off = [], 0
    for shp in _PARAM_SHAPES:
        n = 1
        for s in shp:
            n *= s
        out.append(flat[off:off + n].view(shp))
        off += n
    assert off == L.NUM_PARAMS
    return out


class NeRFMLP(torch.autograd.Function):
    """raw[M,4] = (rgb pre-sigmoid, sigma pre-ReLU).  dirs has one row per `group` consecutive points."""

    @staticmethod
    def forward(ctx, pts, dirs, group, cache: PackedWeights, grad_mode: bool, *params):
        pts, dirs = _f32(pts, "x"), _f32(dirs, "d")
        for i, p in enumerate(params):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"NeRF parameter {i} must be a contiguous fp32 CUDA tensor (no CPU fallback)")
        M = pts.shape[0]
        if M % group != 0 or dirs.shape[0] * group != M:
            raise ValueError("dirs must have M/group rows")
        packed = cache.get(params)
        # grad_mode = torch.is_grad_enabled() sampled by the caller: inside Function.forward it is always
        # False, and ctx.needs_input_grad ignores no_grad()
        training = bool(grad_mode) and any(ctx.needs_input_grad)
        raw = _empty((M, 4), pts)
        if M == 0:
            ctx.meta = None
            return raw
        nbytes = (L.lib().rn_mlp_workspace_bytes(M, 1) if training else L.lib().rn_mlp_infer_workspace_bytes(M, int(group)))
        ws = torch.empty(nbytes, device=pts.device, dtype=torch.uint8)
        call("rn_mlp_fwd", ptr(packed), ptr(pts), ptr(dirs), M, int(group), ptr(ws), int(training), ptr(raw), stream_ptr())
        if training:
            ctx.save_for_backward(pts, dirs)
            ctx.ws, ctx.packed, ctx.cache = ws, packed, cache
        ctx.meta = (M, int(group), training)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        n_in = 5
        if ctx.meta is None:
            return (None,) * (n_in + L.NUM_PARAM_TENSORS)
        M, group, training = ctx.meta
        pts, dirs = ctx.saved_tensors
        g_raw = _f32(g_raw, "g_raw")
        sink = ctx.cache.grad_sink
        accumulate = sink is not None and ctx.cache.sink_dirty
        flat = sink if (sink is not None and not accumulate) else _empty((L.NUM_PARAMS,), pts)
        g_pts = _empty((M, 3), pts) if ctx.needs_input_grad[0] else None
        g_dirs = _empty((M // group, 3), pts) if ctx.needs_input_grad[1] else None
        call("rn_mlp_bwd", ptr(ctx.packed), ptr(pts), ptr(dirs), M, group, ptr(ctx.ws), ptr(g_raw), ptr(flat), ptr(g_pts),
             ptr(g_dirs), stream_ptr())
        ctx.ws = None
        if sink is not None:
            if accumulate:
                sink.add_(flat)
            ctx.cache.sink_dirty = True
            if ctx.cache.on_grads_ready is not None and not accumulate:
                ctx.cache.on_grads_ready()
            return (g_pts, g_dirs, None, None, None) + (None,) * L.NUM_PARAM_TENSORS
        grads = split_flat_grads(flat)
        grads = [g if ctx.needs_input_grad[n_in + i] else None for i, g in enumerate(grads)]
        return (g_pts, g_dirs, None, None, None, *grads)


def view_dirs(rays_d):
    """rays_d / |rays_d| (rendering.py:165) in one launch; no autograd (callers with a gradient use the aten expression)."""
    rd = _f32(rays_d, "rays_d")
    out = _empty(rd.shape, rd)
    call("rn_view_dirs", ptr(rd), rd.numel() // 3, ptr(out), stream_ptr())
    return out


_render_ws = {}


def render_view(model_coarse, model_fine, pose, H, W, focal, z_base, u_det, white_background=True, ray_begin=0, ray_end=None,
                tile_rays=131072, tile_first=0, tile_step=1, rays_o=None, rays_d=None, out=None, want_depth_acc=True):
    """Evaluation render of (a tile-sharded part of) one view in ONE C-ABI call (`rn_render_view`): rays generated in-kernel
    from `pose` (or given), coarse -> deterministic resampling -> fine, nine launches per tile, no Python tile loop.
    Returns (rgb [n,3], depth [n] | None, acc [n] | None, rays_rendered); rows of tiles this call does not own are left as
    they are in `out` (zeros when allocated here)."""
    dev = (pose if pose is not None else rays_o).device
    pc = model_coarse._packed.get(model_coarse.kernel_params())
    pf = None if model_fine is None else model_fine._packed.get(model_fine.kernel_params())
    Nc, Nf = int(z_base.numel()), (0 if (pf is None or u_det is None) else int(u_det.numel()))
    if ray_end is None:
        ray_end = H * W if pose is not None else rays_o.shape[0]
    n = int(ray_end - ray_begin)
    rgb = out if out is not None else torch.zeros((n, 3), device=dev)
    depth = torch.zeros((n,), device=dev) if want_depth_acc else None
    acc = torch.zeros((n,), device=dev) if want_depth_acc else None
    tile = int(min(tile_rays, max(n, 1)))
    key = (dev, tile, Nc, Nf)
    ws = _render_ws.get(key)
    if ws is None:
        _render_ws.clear()                      # one cached workspace per device/shape: a different shape replaces it
        ws = _render_ws[key] = torch.empty(L.lib().rn_render_workspace_bytes(tile, Nc, Nf), device=dev, dtype=torch.uint8)
    done = ctypes.c_int64(0)
    p = None if pose is None else _f32(pose, "pose")
    ro = None if rays_o is None else _f32(rays_o, "rays_o")
    rd = None if rays_d is None else _f32(rays_d, "rays_d")
    call("rn_render_view", ptr(pc), ptr(pf), ptr(p), ptr(ro), ptr(rd), int(H), int(W), float(focal), W / 2.0, H / 2.0,
         int(ray_begin), int(ray_end), tile, int(tile_first), int(tile_step), ptr(_f32(z_base, "z_base")), Nc,
         ptr(None if u_det is None else _f32(u_det, "u_det")), Nf, int(white_background), ptr(ws), ptr(rgb), ptr(depth), ptr(acc),
         ctypes.byref(done), stream_ptr())
    return rgb, depth, acc, int(done.value)


class HeadAct(torch.autograd.Function):
    """rgb = sigmoid(raw[:, :3]), sigma = relu(raw[:, 3:4])   (model.py:181,194)."""

    @staticmethod
    def forward(ctx, raw):
        raw = _f32(raw, "raw")
        M = raw.shape[0]
        rgb, sigma = _empty((M, 3), raw), _empty((M, 1), raw)
        call("rn_head_act_fwd", ptr(raw), M, ptr(rgb), ptr(sigma), stream_ptr())
        ctx.save_for_backward(raw)
        return rgb, sigma

    @staticmethod
    def backward(ctx, g_rgb, g_sigma):
        (raw,) = ctx.saved_tensors
        M = raw.shape[0]
        g_rgb = None if g_rgb is None else _f32(g_rgb, "g_rgb")
        g_sigma = None if g_sigma is None else _f32(g_sigma, "g_sigma")
        g_raw = _empty((M, 4), raw)
        call("rn_head_act_bwd", ptr(raw), M, ptr(g_rgb), ptr(g_sigma), ptr(g_raw), stream_ptr())
        return g_raw


class PosEnc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, num_freqs):
        x = _f32(x, "x")
        C = x.shape[-1]
        n = x.numel() // C
        out = _empty((*x.shape[:-1], C * (1 + 2 * num_freqs)), x)
        call("rn_posenc_fwd", ptr(x), n, C, int(num_freqs), ptr(out), stream_ptr())
        ctx.save_for_backward(x)
        ctx.L = int(num_freqs)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        C = x.shape[-1]
        g = _f32(g, "g_out")
        gx = _empty(x.shape, x)
        call("rn_posenc_bwd", ptr(x), x.numel() // C, C, ctx.L, ptr(g), ptr(gx), stream_ptr())
        return gx, None


# --------------------------------------------------------------------------------------------
# building block for tests / profiling
# --------------------------------------------------------------------------------------------
_gemm_scratch = {}


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """bool/real [M,N] (N % 32 == 0) -> packed uint32-as-int32 [M, N/32]; column 32c + 2i + h of row m sits in word c at
    bit i (h = 0) or 16 + i (h = 1)  (csrc/mlp_layout.h: relu_mask_bit; test helper, the kernels produce / consume
    this layout)."""
    M, N = mask.shape
    b = (mask > 0).reshape(M, N // 32, 32).to(torch.int64)
    j = torch.arange(32, device=mask.device, dtype=torch.int64)
    pos = torch.where(j % 2 == 1, 16 + j // 2, j // 2)
    w = (b << pos).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()


def gemm_bf16(mode, A, B, bias=None, relu=False, mask_bits=None, out=None, want_mask=False):
    """mode 0: A[M,K] @ B[N,K]^T (+bias, relu) -> bf16 [M,N] (and, with want_mask, the packed ReLU mask of the output);
    mode 1: (A[M,K] @ B[K,N]) with packed mask_bits applied -> bf16; mode 2: A[K,Mo]^T @ B[K,N] -> (fp32 [Mo,N], colsum [Mo])."""
    for name, t in (("A", A), ("B", B), ("out", out)):
        if t is not None and (not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1):
            raise RuntimeError(f"{name} must be a 2-D bf16 CUDA tensor with unit inner stride")
    dev = A.device
    if dev not in _gemm_scratch:
        _gemm_scratch[dev] = torch.empty(L.lib().rn_gemm_scratch_bytes(), device=dev, dtype=torch.uint8)
    sc = _gemm_scratch[dev]
    if mode == 0:
        M, K = A.shape
        N = B.shape[0]
        D = torch.empty((M, N), device=dev, dtype=torch.bfloat16) if out is None else out
        mo = torch.empty((M, N // 32), device=dev, dtype=torch.int32) if want_mask else None
        call("rn_gemm_bf16", 0, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(D), D.stride(0), M, N, K, ptr(bias), int(relu),
             None, ptr(mo), None, ptr(sc), sc.numel(), stream_ptr())
        return (D, mo) if want_mask else D
    if mode == 1:
        M, K = A.shape
        N = B.shape[1]
        D = torch.empty((M, N), device=dev, dtype=torch.bfloat16) if out is None else out
        if mask_bits is not None:
            mask_bits = require_cuda(mask_bits, "mask_bits", torch.int32)
        call("rn_gemm_bf16", 1, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(D), D.stride(0), M, N, K, None, 0,
             ptr(mask_bits), None, None, ptr(sc), sc.numel(), stream_ptr())
        return D
    K, Mo = A.shape
    N = B.shape[1]
    D = torch.empty((Mo, N), device=dev, dtype=torch.float32)
    colsum = torch.empty((Mo,), device=dev, dtype=torch.float32)
    call("rn_gemm_bf16", 2, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(D), D.stride(0), Mo, N, K, None, 0, None, None,
         ptr(colsum), ptr(sc), sc.numel(), stream_ptr())
    return D, colsum
