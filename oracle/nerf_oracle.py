"""CPU oracle for the Robust-NeRF render-and-train hot path (numpy, fp32).

TEST INFRASTRUCTURE ONLY.  This module restates, in plain numpy, the algorithm of
the reference's PyTorch hot path so the CUDA kernels can be checked on a box where
`/root/reference` does not exist.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
package (`robust-nerf_b200/`) never does.

Pinning: every function here is checked in `tests/test_oracle_golden.py` against
vectors produced by importing and running the UNMODIFIED reference in the authoring
container (`tests/golden/make_golden.py`, outputs in `tests/golden/*.npz`), and
against the reference's only recorded known-answer data
(`outputs/*/final_poses.pt` pose errors).

Each function cites the reference file:line it follows (paths relative to the
reference repo root).  Arithmetic is float32 end to end, like the reference.
Summation-order contract for the inverse-CDF sampler (`sample_pdf`): the weight sum and
the cumulative sum follow a fixed 32-lane scan order (`_warp_scan`: contiguous chunk per
lane summed left to right, Kogge-Stone over the lane totals) -- the order the CUDA kernel
implements (`warp_build_cdf`), so sample indices are bit-exact between the two.
Golden vectors of the later rows: `tests/golden/make_golden_metrics.py` (metrics.py) and
`tests/golden/make_golden_noise.py` (noise.py), checked in `tests/test_oracle_golden.py` /
`tests/test_host_cpu.py`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# config mirrors (noisy_src/config.py:10-43)
# --------------------------------------------------------------------------------------
@dataclass
class ModelConfig:
    pos_freqs: int = 10
    dir_freqs: int = 4
    hidden_dim: int = 256
    num_hidden_layers: int = 8
    skips: Tuple[int, ...] = (4,)
    use_view_dirs: bool = True


@dataclass
class RenderConfig:
    near: float = 2.0
    far: float = 6.0
    num_samples: int = 64
    num_samples_fine: int = 128
    use_hierarchical: bool = True
    perturb: bool = True
    raw_noise_std: float = 0.0
    white_background: bool = True


def _f32(x) -> np.ndarray:
    return np.asarray(x, dtype=F32)


def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace semantics in fp32 (symmetric two-sided evaluation).

    torch computes step=(end-start)/(steps-1) in fp32, the lower half as
    fma(step, i, start) and the upper half as fma(-step, steps-1-i, end); used by
    noisy_src/rays.py:185 and :252.
    """
    if steps == 1:
        return np.array([start], dtype=F32)
    start32, end32 = F32(start), F32(end)
    step = F32((end32 - start32) / F32(steps - 1))
    idx = np.arange(steps)
    half = steps // 2
    # torch's CPU kernel evaluates start + step*i / end - step*(n-1-i) with a fused
    # multiply-add (one rounding); emulate it exactly through float64.
    lo = (np.float64(start32) + np.float64(step) * idx).astype(F32)
    hi = (np.float64(end32) - np.float64(step) * (steps - 1 - idx)).astype(F32)
    return np.where(idx < half, lo, hi).astype(F32)


# --------------------------------------------------------------------------------------
# rays.py
# --------------------------------------------------------------------------------------
def get_ray_directions(H: int, W: int, focal: float, center=None) -> np.ndarray:
    """noisy_src/rays.py:17-64.  No half-pixel offset; -Z forward, +Y up."""
    if center is None:
        cx, cy = W / 2.0, H / 2.0
    else:
        cx, cy = center
    i = np.broadcast_to(np.arange(W, dtype=F32)[None, :], (H, W))
    j = np.broadcast_to(np.arange(H, dtype=F32)[:, None], (H, W))
    f = F32(focal)
    dx = ((i - F32(cx)) / f).astype(F32)
    dy = (-((j - F32(cy)) / f)).astype(F32)
    dz = -np.ones_like(dx)
    return np.stack([dx, dy, dz], axis=-1).astype(F32)


def get_rays(directions: np.ndarray, c2w: np.ndarray):
    """noisy_src/rays.py:67-99.  rays_d = normalise(R @ dir), rays_o = t."""
    directions = _f32(directions)
    R = _f32(c2w)[:3, :3]
    prod = directions[..., None, :] * R  # (...,3,3)
    rays_d = ((prod[..., 0] + prod[..., 1]) + prod[..., 2]).astype(F32)
    nrm = np.sqrt(((rays_d[..., 0] * rays_d[..., 0] + rays_d[..., 1] * rays_d[..., 1])
                   + rays_d[..., 2] * rays_d[..., 2]).astype(F32)).astype(F32)
    rays_d = (rays_d / nrm[..., None]).astype(F32)
    rays_o = np.broadcast_to(_f32(c2w)[:3, 3], rays_d.shape).copy()
    return rays_o, rays_d


def get_rays_batch(H: int, W: int, focal: float, c2w_batch: np.ndarray):
    """noisy_src/rays.py:102-142."""
    dirs = get_ray_directions(H, W, focal)
    o, d = zip(*[get_rays(dirs, c) for c in c2w_batch])
    return np.stack(o, 0), np.stack(d, 0)


def stratified_z(near: float, far: float, num_samples: int, batch_shape, perturb: bool,
                 lindisp: bool = False, t_rand: Optional[np.ndarray] = None) -> np.ndarray:
    """z-values of noisy_src/rays.py:185-205."""
    t = linspace_f32(0.0, 1.0, num_samples)
    if lindisp:
        z = (F32(1.0) / (F32(1.0 / near) * (F32(1.0) - t) + F32(1.0 / far) * t)).astype(F32)
    else:
        z = (F32(near) * (F32(1.0) - t) + F32(far) * t).astype(F32)
    z = np.broadcast_to(z, (*batch_shape, num_samples)).astype(F32)
    if perturb:
        assert t_rand is not None, "oracle needs the uniform draws explicitly"
        mids = (F32(0.5) * (z[..., 1:] + z[..., :-1])).astype(F32)
        upper = np.concatenate([mids, z[..., -1:]], -1)
        lower = np.concatenate([z[..., :1], mids], -1)
        z = (lower + (upper - lower) * _f32(t_rand)).astype(F32)
    return z


def sample_along_rays(rays_o, rays_d, near, far, num_samples, perturb=True, lindisp=False,
                      t_rand=None):
    """noisy_src/rays.py:145-210."""
    rays_o, rays_d = _f32(rays_o), _f32(rays_d)
    z = stratified_z(near, far, num_samples, rays_o.shape[:-1], perturb, lindisp, t_rand)
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]).astype(F32)
    return pts, z


def _warp_scan(x: np.ndarray):
    """Inclusive fp32 prefix sums over the last axis in the order a 32-lane warp scan produces them, and the total.

    torch's own summation order is backend-dependent (pairwise / vectorised sum and a sequential cumsum on the CPU,
    a parallel block scan on CUDA), so the oracle has to fix one; it fixes the order of the CUDA kernel
    (csrc/sampling.cu: warp_build_cdf): with C = ceil(n / 32), lane l owns the contiguous chunk [l*C, (l+1)*C) and sums it
    left to right; the 32 lane totals go through a Kogge-Stone scan (offsets 1, 2, 4, 8, 16: v[l] += v[l - o] for
    l >= o, all lanes at once); element k = l*C + j gets offset[l] + local_prefix[j], offset[0] = 0.
    """
    x = np.asarray(x, dtype=F32)
    n = x.shape[-1]
    C = (n + 31) // 32
    pad = np.zeros((*x.shape[:-1], 32 * C - n), dtype=F32)
    xp = np.concatenate([x, pad], -1).reshape(*x.shape[:-1], 32, C)
    local = np.empty_like(xp)
    run = np.zeros(xp.shape[:-1], dtype=F32)
    for j in range(C):
        run = (run + xp[..., j]).astype(F32)
        local[..., j] = run
    v = run.copy()                                              # lane totals
    for o in (1, 2, 4, 8, 16):
        shifted = np.zeros_like(v)
        shifted[..., o:] = v[..., :-o]
        v = np.where(np.arange(32) >= o, (v + shifted).astype(F32), v).astype(F32)
    excl = np.zeros_like(v)
    excl[..., 1:] = v[..., :-1]
    out = (excl[..., None] + local).astype(F32).reshape(*x.shape[:-1], 32 * C)[..., :n]
    return out, v[..., 31].copy()


def _seq_sum(x: np.ndarray) -> np.ndarray:
    """Total of the last axis in the warp-scan order (name kept: it is the kernel's order)."""
    return _warp_scan(x)[1]


def _seq_cumsum(x: np.ndarray) -> np.ndarray:
    return _warp_scan(x)[0]


def sample_pdf(bins, weights, num_samples, det=False, u=None, return_inds=False):
    """noisy_src/rays.py:213-279 (inverse-CDF sampling, searchsorted right=True)."""
    bins, weights = _f32(bins), _f32(weights)
    w = (weights + F32(1e-5)).astype(F32)
    pdf = (w / _seq_sum(w)[..., None]).astype(F32)
    cdf = np.concatenate([np.zeros_like(pdf[..., :1]), _seq_cumsum(pdf)], -1)  # (..., nb)
    nb = cdf.shape[-1]
    if det:
        u = np.broadcast_to(linspace_f32(0.0, 1.0, num_samples), (*cdf.shape[:-1], num_samples))
    else:
        assert u is not None, "oracle needs the uniform draws explicitly"
    u = _f32(u)
    # searchsorted(right=True): number of cdf entries <= u
    inds = (cdf[..., None, :] <= u[..., :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, nb - 1)
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bin_b = np.take_along_axis(bins, below, -1)
    bin_a = np.take_along_axis(bins, above, -1)
    denom = (cdf_a - cdf_b).astype(F32)
    denom = np.where(denom < F32(1e-5), F32(1.0), denom).astype(F32)
    t = ((u - cdf_b) / denom).astype(F32)
    samples = (bin_b + t * (bin_a - bin_b)).astype(F32)
    if return_inds:
        return samples, inds, cdf
    return samples


def sample_hierarchical(rays_o, rays_d, z_vals, weights, num_samples_fine, det=False, u=None,
                        return_inds=False):
    """noisy_src/rays.py:282-333: mid-point bins, interior weights, sorted merge."""
    rays_o, rays_d, z_vals, weights = map(_f32, (rays_o, rays_d, z_vals, weights))
    mids = (F32(0.5) * (z_vals[..., 1:] + z_vals[..., :-1])).astype(F32)
    res = sample_pdf(mids, weights[..., 1:-1], num_samples_fine, det=det, u=u,
                     return_inds=return_inds)
    z_samples = res[0] if return_inds else res
    z_all = np.sort(np.concatenate([z_vals, z_samples], -1), axis=-1).astype(F32)
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]).astype(F32)
    if return_inds:
        return pts, z_all, res[1]
    return pts, z_all


# --------------------------------------------------------------------------------------
# model.py
# --------------------------------------------------------------------------------------
def positional_encoding(x: np.ndarray, num_freqs: int) -> np.ndarray:
    """noisy_src/model.py:58-80: [x, sin(2^k x), cos(2^k x)]_k, NO pi factor."""
    x = _f32(x)
    out = [x]
    for k in range(num_freqs):
        f = F32(2.0 ** k)
        out.append(np.sin(f * x).astype(F32))
        out.append(np.cos(f * x).astype(F32))
    return np.concatenate(out, -1).astype(F32)


def positional_encoding_fast(x: np.ndarray, num_freqs: int) -> np.ndarray:
    """The features as the CUDA path builds its bf16 operand (csrc/pe.cuh): the argument is reduced once, in turns --
    t = x / (2 pi) as a two-float value (FMA residual), r_k = frac(2^k t_hi) + 2^k t_lo, angle = 2 pi r_k in fp32 -- and
    the hardware's approximate sin / cos of the reduced angle (absolute error 2^-21.4) is modelled by the exact one.
    Within 8e-7 of `positional_encoding` at every frequency; used by the `emulate_bf16` mode only."""
    x = _f32(x)
    c = 1.0 / (2.0 * np.pi)
    c_hi = F32(c)
    c_lo = F32(c - float(c_hi))
    t_hi = (x * c_hi).astype(F32)
    resid = (x.astype(np.float64) * float(c_hi) - t_hi.astype(np.float64)).astype(F32)      # fma(x, c_hi, -t_hi): exact
    t_lo = (x.astype(np.float64) * float(c_lo) + resid.astype(np.float64)).astype(F32)      # fma(x, c_lo, resid)
    out = [x]
    for k in range(num_freqs):
        f = F32(2.0 ** k)
        y = (t_hi * f).astype(F32)
        d = (y - np.rint(y)).astype(F32)
        r = (t_lo.astype(np.float64) * float(f) + d.astype(np.float64)).astype(F32)          # fma(t_lo, f, d)
        ang = (r * F32(6.2831855)).astype(F32)
        out.append(np.sin(ang.astype(np.float64)).astype(F32))
        out.append(np.cos(ang.astype(np.float64)).astype(F32))
    return np.concatenate(out, -1).astype(F32)


def positional_encoding_backward(x: np.ndarray, num_freqs: int, g: np.ndarray) -> np.ndarray:
    x, g = _f32(x), _f32(g)
    C = x.shape[-1]
    dx = g[..., :C].copy()
    for k in range(num_freqs):
        f = F32(2.0 ** k)
        gs = g[..., C * (1 + 2 * k): C * (2 + 2 * k)]
        gc = g[..., C * (2 + 2 * k): C * (3 + 2 * k)]
        dx += f * (np.cos(f * x) * gs - np.sin(f * x) * gc)
    return dx.astype(F32)


def param_names(cfg: ModelConfig):
    names = []
    for i in range(cfg.num_hidden_layers):
        names += [f"pts_linears.{i}.weight", f"pts_linears.{i}.bias"]
    for n in ("sigma_linear", "feature_linear", "dir_linear", "rgb_linear"):
        names += [f"{n}.weight", f"{n}.bias"]
    return names


def param_shapes(cfg: ModelConfig) -> Dict[str, Tuple[int, ...]]:
    """Shapes of noisy_src/model.py:104-143 (weight is (out, in))."""
    pos_dim = 3 * (1 + 2 * cfg.pos_freqs)
    dir_dim = 3 * (1 + 2 * cfg.dir_freqs)
    shapes = {}
    in_dim = pos_dim
    for i in range(cfg.num_hidden_layers):
        shapes[f"pts_linears.{i}.weight"] = (cfg.hidden_dim, in_dim)
        shapes[f"pts_linears.{i}.bias"] = (cfg.hidden_dim,)
        in_dim = cfg.hidden_dim + (pos_dim if i in cfg.skips else 0)
    shapes["sigma_linear.weight"] = (1, cfg.hidden_dim)
    shapes["sigma_linear.bias"] = (1,)
    shapes["feature_linear.weight"] = (cfg.hidden_dim, cfg.hidden_dim)
    shapes["feature_linear.bias"] = (cfg.hidden_dim,)
    d_in = cfg.hidden_dim + (dir_dim if cfg.use_view_dirs else 0)
    shapes["dir_linear.weight"] = (cfg.hidden_dim // 2, d_in)
    shapes["dir_linear.bias"] = (cfg.hidden_dim // 2,)
    shapes["rgb_linear.weight"] = (3, cfg.hidden_dim // 2)
    shapes["rgb_linear.bias"] = (3,)
    return shapes


def make_weights(seed: int, cfg: Optional[ModelConfig] = None, sharpen: bool = False
                 ) -> Dict[str, np.ndarray]:
    """Deterministic numpy weight generator, distributed like nn.Linear's default init
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias).  `sharpen` applies the
    SURVEY section 7 hard-part 6 transform so density is large enough to exercise
    compositing, CDF inversion and the saturated-alpha paths."""
    cfg = cfg or ModelConfig()
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        lname = name.rsplit(".", 1)[0] + ".weight"
        fan_in = param_shapes(cfg)[lname][1]
        bound = 1.0 / np.sqrt(fan_in)
        out[name] = rng.uniform(-bound, bound, size=shape).astype(F32)
    if sharpen:
        out["sigma_linear.weight"] = (out["sigma_linear.weight"] * F32(300.0)).astype(F32)
        out["sigma_linear.bias"] = (out["sigma_linear.bias"] + F32(0.5)).astype(F32)
        out["rgb_linear.weight"] = (out["rgb_linear.weight"] * F32(30.0)).astype(F32)
    return out


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bfloat16 (ties to even), returned as fp32.  Used by the
    `emulate_bf16` mode, which restates WHERE the CUDA path rounds (weights, PE features, every
    stored activation and back-propagated gradient) so ReLU masks coincide and the kernels can be
    checked tightly; accumulation stays fp32 like the tensor cores'."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(F32).reshape(np.shape(x))


def _sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


# Which rounding points of the CUDA path `emulate_bf16` restates.  All on = the kernels as built.  The ablation in
# scripts/grad_rounding_ablation.py switches groups off to measure where the distance to the fp32 reference comes from:
# "w" weights, "act" stored activations, "enc" positional-encoding features (forward operands); "dh" back-propagated
# activation gradients, "dxe" the three 64-wide encoding-gradient outputs, "bw" weights as read by the data-gradient GEMMs.
BF16_POINTS = {"w": True, "act": True, "enc": True, "dh": True, "dxe": True, "bw": True}


def _rounders(emulate: bool):
    ident = (lambda a: a)
    return {k: (bf16_round if (emulate and on) else ident) for k, on in BF16_POINTS.items()}


def nerf_forward(params: Dict[str, np.ndarray], x, d, cfg: Optional[ModelConfig] = None,
                 keep_cache: bool = False, emulate_bf16: bool = False):
    """noisy_src/model.py:145-196.  Returns rgb (N,3), sigma (N,1)[, cache]."""
    cfg = cfg or ModelConfig()
    rr = _rounders(emulate_bf16)
    rW, rA, rE = rr["w"], rr["act"], rr["enc"]
    x = _f32(x)
    pe = positional_encoding_fast if (emulate_bf16 and BF16_POINTS["enc"]) else positional_encoding
    x_enc = rE(pe(x, cfg.pos_freqs))
    h = x_enc
    cache = {"x": x, "x_enc": x_enc, "ins": [], "pre": [], "emulate": emulate_bf16}
    for i in range(cfg.num_hidden_layers):
        W, b = rW(params[f"pts_linears.{i}.weight"]), params[f"pts_linears.{i}.bias"]
        cache["ins"].append(h)
        pre = (h @ W.T + b).astype(F32)
        h = rA(np.maximum(pre, F32(0.0)))
        cache["pre"].append(h if emulate_bf16 else pre)   # mask = stored activation > 0
        if i in cfg.skips:
            h = np.concatenate([x_enc, h], -1)
    sig_pre = (h @ params["sigma_linear.weight"].T + params["sigma_linear.bias"]).astype(F32)
    sigma = np.maximum(sig_pre, F32(0.0))
    feats = rA((h @ rW(params["feature_linear.weight"]).T + params["feature_linear.bias"]).astype(F32))
    if cfg.use_view_dirs:
        if d is None:
            raise ValueError("reference crashes for d=None with use_view_dirs (model.py:187-193)")
        d = _f32(d)
        d_enc = rE(pe(d, cfg.dir_freqs))
        hc_in = np.concatenate([feats, d_enc], -1)
    else:
        d_enc, hc_in = None, feats
    hc_pre = (hc_in @ rW(params["dir_linear.weight"]).T + params["dir_linear.bias"]).astype(F32)
    hc = rA(np.maximum(hc_pre, F32(0.0)))
    rgb_pre = (hc @ params["rgb_linear.weight"].T + params["rgb_linear.bias"]).astype(F32)
    rgb = _sigmoid(rgb_pre)
    if keep_cache:
        cache.update(h_last=h, sig_pre=sig_pre, d=d, hc_in=hc_in, hc_pre=hc if emulate_bf16 else hc_pre, hc=hc, rgb=rgb)
        return rgb, sigma, cache
    return rgb, sigma


def nerf_backward(params, cache, g_rgb, g_sigma, cfg: Optional[ModelConfig] = None,
                  need_input_grad: bool = False):
    """Manual backward of nerf_forward.  Returns (param grads dict, dx, dd)."""
    cfg = cfg or ModelConfig()
    rr = _rounders(bool(cache.get("emulate")))
    rh, rx, rw = rr["dh"], rr["dxe"], rr["bw"]
    g = {}
    rgb = cache["rgb"]
    g_rgb_pre = (_f32(g_rgb) * rgb * (F32(1.0) - rgb)).astype(F32)
    g["rgb_linear.weight"] = (g_rgb_pre.T @ cache["hc"]).astype(F32)
    g["rgb_linear.bias"] = g_rgb_pre.sum(0).astype(F32)
    g_hc = (g_rgb_pre @ params["rgb_linear.weight"]).astype(F32)
    g_hc_pre = rh((g_hc * (cache["hc_pre"] > 0)).astype(F32))
    g["dir_linear.weight"] = (g_hc_pre.T @ cache["hc_in"]).astype(F32)
    g["dir_linear.bias"] = g_hc_pre.sum(0).astype(F32)
    H = cfg.hidden_dim
    g_hc_in = (g_hc_pre @ rw(params["dir_linear.weight"])).astype(F32)
    g_hc_in = np.concatenate([rh(g_hc_in[:, :H]), rx(g_hc_in[:, H:])], -1)
    g_feats = g_hc_in[:, :H]
    g_d_enc = g_hc_in[:, H:] if cfg.use_view_dirs else None
    h_last = cache["h_last"]
    g["feature_linear.weight"] = (g_feats.T @ h_last).astype(F32)
    g["feature_linear.bias"] = g_feats.sum(0).astype(F32)
    g_sig_pre = rh((_f32(g_sigma).reshape(-1, 1) * (cache["sig_pre"] > 0)).astype(F32))
    g["sigma_linear.weight"] = (g_sig_pre.T @ h_last).astype(F32)
    g["sigma_linear.bias"] = g_sig_pre.sum(0).astype(F32)
    g_h = (g_feats @ rw(params["feature_linear.weight"]) + g_sig_pre @ rw(params["sigma_linear.weight"])).astype(F32)
    pos_dim = cache["x_enc"].shape[-1]
    g_x_enc = np.zeros_like(cache["x_enc"])
    for i in reversed(range(cfg.num_hidden_layers)):
        if i in cfg.skips:
            g_x_enc += rx(g_h[:, :pos_dim])
            g_h = g_h[:, pos_dim:]
        g_pre = rh((g_h * (cache["pre"][i] > 0)).astype(F32))
        g[f"pts_linears.{i}.weight"] = (g_pre.T @ cache["ins"][i]).astype(F32)
        g[f"pts_linears.{i}.bias"] = g_pre.sum(0).astype(F32)
        if i > 0 or need_input_grad:
            g_h = (g_pre @ rw(params[f"pts_linears.{i}.weight"])).astype(F32)
    dx = dd = None
    if need_input_grad:
        g_x_enc += rx(g_h)
        dx = positional_encoding_backward(cache["x"], cfg.pos_freqs, g_x_enc)
        if cfg.use_view_dirs:
            dd = positional_encoding_backward(cache["d"], cfg.dir_freqs, g_d_enc)
    return g, dx, dd


# --------------------------------------------------------------------------------------
# rendering.py
# --------------------------------------------------------------------------------------
def raw2outputs(rgb, sigma, z_vals, rays_d, noise=None, white_background=True, keep_cache=False):
    """noisy_src/rendering.py:20-116.  `noise` = the already-scaled randn draw (or None)."""
    rgb, sigma, z_vals, rays_d = map(_f32, (rgb, sigma, z_vals, rays_d))
    sigma = sigma.reshape(z_vals.shape)
    dz = np.concatenate([z_vals[..., 1:] - z_vals[..., :-1],
                         np.full_like(z_vals[..., :1], 1e10)], -1).astype(F32)
    nrm = np.sqrt(((rays_d[..., 0] ** 2 + rays_d[..., 1] ** 2) + rays_d[..., 2] ** 2)).astype(F32)
    dists = (dz * nrm[..., None]).astype(F32)
    if noise is not None:
        sigma = (sigma + _f32(noise)).astype(F32)
    sr = np.maximum(sigma, F32(0.0))
    alpha = (F32(1.0) - np.exp(-sr * dists)).astype(F32)
    t = (F32(1.0) - alpha + F32(1e-10)).astype(F32)
    T = np.ones_like(alpha)
    for s in range(1, alpha.shape[-1]):
        T[..., s] = (T[..., s - 1] * t[..., s - 1]).astype(F32)
    w = (alpha * T).astype(F32)
    rgb_map = (w[..., None] * rgb).sum(-2).astype(F32)
    depth_map = (w * z_vals).sum(-1).astype(F32)
    acc_map = w.sum(-1).astype(F32)
    if white_background:
        rgb_map = (rgb_map + (F32(1.0) - acc_map[..., None])).astype(F32)
    out = {"rgb_map": rgb_map, "depth_map": depth_map, "acc_map": acc_map, "weights": w}
    if keep_cache:
        out["_cache"] = dict(rgb=rgb, sigma=sigma, z=z_vals, rays_d=rays_d, dz=dz, nrm=nrm,
                             dists=dists, alpha=alpha, t=t, T=T, w=w, white=white_background)
    return out


def raw2outputs_backward(cache, g_rgb_map, g_depth=None, g_acc=None, g_weights=None):
    """Analytic backward of raw2outputs.  Returns d_rgb (..,S,3), d_sigma (..,S), d_rays_d."""
    c = cache
    rgb, z, w, T, t, alpha = c["rgb"], c["z"], c["w"], c["T"], c["t"], c["alpha"]
    g_rgb_map = _f32(g_rgb_map)
    G = (g_rgb_map[..., None, :] * rgb).sum(-1).astype(F32)
    if c["white"]:
        G = G - g_rgb_map.sum(-1)[..., None]
    if g_depth is not None:
        G = G + _f32(g_depth)[..., None] * z
    if g_acc is not None:
        G = G + _f32(g_acc)[..., None]
    if g_weights is not None:
        G = G + _f32(g_weights)
    G = G.astype(F32)
    d_rgb = (w[..., None] * g_rgb_map[..., None, :]).astype(F32)
    Gw = (G * w).astype(F32)
    # suffix sums R_s = sum_{k>s} G_k w_k
    R = np.zeros_like(Gw)
    for s in range(Gw.shape[-1] - 2, -1, -1):
        R[..., s] = (R[..., s + 1] + Gw[..., s + 1]).astype(F32)
    d_alpha = (G * T - R / t).astype(F32)
    one_m = (F32(1.0) - alpha).astype(F32)
    pos = c["sigma"] > 0
    d_sigma = (d_alpha * c["dists"] * one_m * pos).astype(F32)
    d_dists = (d_alpha * np.maximum(c["sigma"], F32(0.0)) * one_m).astype(F32)
    d_nrm = (d_dists * c["dz"]).sum(-1).astype(F32)
    d_rays_d = (d_nrm[..., None] * c["rays_d"] / c["nrm"][..., None]).astype(F32)
    return d_rgb, d_sigma, d_rays_d


def render_rays(params_c, params_f, rays_o, rays_d, rcfg: Optional[RenderConfig] = None,
                mcfg: Optional[ModelConfig] = None, is_train=True, t_rand=None, u=None,
                noise_c=None, noise_f=None, keep_cache=False, emulate_bf16=False):
    """noisy_src/rendering.py:119-240."""
    rcfg, mcfg = rcfg or RenderConfig(), mcfg or ModelConfig()
    rays_o, rays_d = _f32(rays_o), _f32(rays_d)
    perturb = rcfg.perturb if is_train else False
    nrm = np.sqrt(((rays_d[:, 0] ** 2 + rays_d[:, 1] ** 2) + rays_d[:, 2] ** 2)).astype(F32)
    viewdirs = (rays_d / nrm[:, None]).astype(F32)
    B, Nc = rays_o.shape[0], rcfg.num_samples
    pts_c, z_c = sample_along_rays(rays_o, rays_d, rcfg.near, rcfg.far, Nc, perturb, t_rand=t_rand)
    vd = np.broadcast_to(viewdirs[:, None, :], (B, Nc, 3)).reshape(-1, 3)
    fc = nerf_forward(params_c, pts_c.reshape(-1, 3), vd, mcfg, keep_cache=keep_cache, emulate_bf16=emulate_bf16)
    out_c = raw2outputs(fc[0].reshape(B, Nc, 3), fc[1].reshape(B, Nc, 1), z_c, rays_d,
                        noise=noise_c if is_train else None,
                        white_background=rcfg.white_background, keep_cache=keep_cache)
    res = {"rgb_coarse": out_c["rgb_map"], "depth_coarse": out_c["depth_map"],
           "acc_coarse": out_c["acc_map"]}
    cache = {"viewdirs": viewdirs, "nrm": nrm, "z_c": z_c, "fc": fc, "out_c": out_c}
    if rcfg.use_hierarchical and params_f is not None:
        pts_f, z_f = sample_hierarchical(rays_o, rays_d, z_c, out_c["weights"],
                                         rcfg.num_samples_fine, det=not is_train, u=u)
        Nt = z_f.shape[-1]
        vd = np.broadcast_to(viewdirs[:, None, :], (B, Nt, 3)).reshape(-1, 3)
        ff = nerf_forward(params_f, pts_f.reshape(-1, 3), vd, mcfg, keep_cache=keep_cache, emulate_bf16=emulate_bf16)
        out_f = raw2outputs(ff[0].reshape(B, Nt, 3), ff[1].reshape(B, Nt, 1), z_f, rays_d,
                            noise=noise_f if is_train else None,
                            white_background=rcfg.white_background, keep_cache=keep_cache)
        res.update(rgb_fine=out_f["rgb_map"], depth_fine=out_f["depth_map"],
                   acc_fine=out_f["acc_map"])
        cache.update(z_f=z_f, ff=ff, out_f=out_f)
    if keep_cache:
        res["_cache"] = cache
        res["z_coarse"], res["weights_coarse"] = z_c, out_c["weights"]
        if "z_f" in cache:
            res["z_fine"] = cache["z_f"]
    return res


def train_step_grads(params_c, params_f, rays_o, rays_d, target, rcfg=None, mcfg=None,
                     t_rand=None, u=None, need_ray_grad=False, emulate_bf16=False):
    """Loss of noisy_src/train.py:88-99 (mse_coarse + mse_fine) and its gradients w.r.t.
    both nets' parameters (and rays_o / rays_d when `need_ray_grad`)."""
    rcfg, mcfg = rcfg or RenderConfig(), mcfg or ModelConfig()
    rays_o, rays_d, target = map(_f32, (rays_o, rays_d, target))
    res = render_rays(params_c, params_f, rays_o, rays_d, rcfg, mcfg, True, t_rand, u,
                      keep_cache=True, emulate_bf16=emulate_bf16)
    cache = res["_cache"]
    B = rays_o.shape[0]
    n = F32(B * 3)
    out = {"rgb_coarse": res["rgb_coarse"]}
    diff_c = res["rgb_coarse"] - target
    loss = F32((diff_c.astype(np.float64) ** 2).mean())
    out["loss_coarse"] = loss
    d_o = np.zeros_like(rays_o)
    d_d = np.zeros_like(rays_d)
    d_vd = np.zeros_like(rays_d)

    def _back(net_params, fwd, outc, z, diff):
        nonlocal d_o, d_d, d_vd
        g_map = (F32(2.0) * diff / n).astype(F32)
        d_rgb, d_sigma, d_rd = raw2outputs_backward(outc["_cache"], g_map)
        S = z.shape[-1]
        g, dx, dd = nerf_backward(net_params, fwd[2], d_rgb.reshape(-1, 3), d_sigma.reshape(-1, 1),
                                  mcfg, need_input_grad=need_ray_grad)
        if need_ray_grad:
            dx = dx.reshape(B, S, 3)
            d_o += dx.sum(1)
            d_d += (dx * z[..., None]).sum(1) + d_rd
            d_vd += dd.reshape(B, S, 3).sum(1)
        return g

    grads_c = _back(params_c, cache["fc"], cache["out_c"], cache["z_c"], diff_c)
    grads_f = None
    if "ff" in cache:
        diff_f = res["rgb_fine"] - target
        out["rgb_fine"] = res["rgb_fine"]
        out["loss_fine"] = F32((diff_f.astype(np.float64) ** 2).mean())
        loss = F32(loss + out["loss_fine"])
        grads_f = _back(params_f, cache["ff"], cache["out_f"], cache["z_f"], diff_f)
    if need_ray_grad:
        # viewdirs = d / |d|  (rendering.py:165)
        vd, nrm = cache["viewdirs"], cache["nrm"]
        d_d += ((d_vd - vd * (vd * d_vd).sum(-1, keepdims=True)) / nrm[:, None]).astype(F32)
        out["d_rays_o"], out["d_rays_d"] = d_o.astype(F32), d_d.astype(F32)
    out.update(loss=loss, grads_coarse=grads_c, grads_fine=grads_f)
    return out


# --------------------------------------------------------------------------------------
# train_pose_opt.py :: CameraPoseParameters, data_pose_opt.py
# --------------------------------------------------------------------------------------
def skew(v: np.ndarray) -> np.ndarray:
    """noisy_src/train_pose_opt.py:165-184."""
    z = np.zeros_like(v[..., 0])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2).astype(F32)


def axis_angle_to_rotation_matrix(w: np.ndarray, keep_cache=False):
    """noisy_src/train_pose_opt.py:122-163 (Rodrigues with the theta<1e-6 -> I branch)."""
    w = _f32(w)
    shp = w.shape[:-1]
    w2 = w.reshape(-1, 3)
    ang = np.sqrt(((w2[:, 0] ** 2 + w2[:, 1] ** 2) + w2[:, 2] ** 2)).astype(F32)[:, None]
    small = ang < F32(1e-6)
    ang = np.where(small, F32(1.0), ang).astype(F32)
    axis = (w2 / ang).astype(F32)
    K = skew(axis)
    K2 = (K @ K).astype(F32)
    I = np.broadcast_to(np.eye(3, dtype=F32), K.shape)
    s, c = np.sin(ang)[..., None].astype(F32), np.cos(ang)[..., None].astype(F32)
    R = (I + s * K + (F32(1.0) - c) * K2).astype(F32)
    R = np.where(small[..., None], I, R).astype(F32)
    if keep_cache:
        return R.reshape(*shp, 3, 3), dict(w=w2, ang=ang, small=small, K=K, K2=K2, s=s, c=c)
    return R.reshape(*shp, 3, 3)


def axis_angle_backward(cache, G: np.ndarray) -> np.ndarray:
    """dL/dw from dL/dR for axis_angle_to_rotation_matrix (zero in the small branch)."""
    G = _f32(G).reshape(-1, 3, 3)
    K, K2, s, c, ang, w = (cache[k] for k in ("K", "K2", "s", "c", "ang", "w"))
    d_s = (G * K).sum((-1, -2))
    d_omc = (G * K2).sum((-1, -2))
    d_ang = d_s * c[:, 0, 0] + d_omc * s[:, 0, 0]
    KT = np.swapaxes(K, -1, -2)
    dK = s * G + (F32(1.0) - c) * (G @ KT + KT @ G)
    dk = np.stack([dK[:, 2, 1] - dK[:, 1, 2], dK[:, 0, 2] - dK[:, 2, 0],
                   dK[:, 1, 0] - dK[:, 0, 1]], -1)
    a = ang[:, 0]
    dw = dk / a[:, None]
    d_ang = d_ang - (dk * w).sum(-1) / (a * a)
    dw = dw + d_ang[:, None] * w / a[:, None]
    dw = np.where(cache["small"], F32(0.0), dw)
    return dw.astype(F32)


def get_poses(initial_poses, rot_deltas, trans_deltas, learn_rotation=True,
              learn_translation=True, keep_cache=False):
    """noisy_src/train_pose_opt.py:186-226: R = exp(w) @ R_init, t = t_init + dt."""
    P0 = _f32(initial_poses)
    cache = {"P0": P0}
    if learn_rotation:
        Rd, rc = axis_angle_to_rotation_matrix(rot_deltas, keep_cache=True)
        R = (Rd @ P0[:, :3, :3]).astype(F32)
        cache["rc"] = rc
    else:
        R = P0[:, :3, :3]
    t = (P0[:, :3, 3] + _f32(trans_deltas)).astype(F32) if learn_translation else P0[:, :3, 3]
    P = np.zeros_like(P0)
    P[:, :3, :3], P[:, :3, 3], P[:, 3, 3] = R, t, F32(1.0)
    return (P, cache) if keep_cache else P


def get_poses_backward(cache, g_poses, learn_rotation=True, learn_translation=True):
    g = _f32(g_poses)
    d_w = d_t = None
    if learn_rotation:
        G = (g[:, :3, :3] @ np.swapaxes(cache["P0"][:, :3, :3], -1, -2)).astype(F32)
        d_w = axis_angle_backward(cache["rc"], G)
    if learn_translation:
        d_t = g[:, :3, 3].copy()
    return d_w, d_t


def pixel_bookkeeping(flat_idx: np.ndarray, H: int, W: int):
    """Index arithmetic equal to the tables of noisy_src/data_pose_opt.py:56-76:
    image = idx // (H*W), v = (idx % (H*W)) // W, u = idx % W  (u, v stored as float32)."""
    flat_idx = np.asarray(flat_idx, dtype=np.int64)
    img = flat_idx // (H * W)
    rem = flat_idx % (H * W)
    v, u = rem // W, rem % W
    return img, np.stack([u, v], -1).astype(F32)


def get_rays_from_pixels(image_indices, pixel_coords, poses, H, W, focal, keep_cache=False):
    """Net effect of noisy_src/data_pose_opt.py:83-148 + 200-223:
    rays_d[b] = normalise(R[img_b] @ dir[v_b,u_b]), rays_o[b] = t[img_b]."""
    img = np.asarray(image_indices, dtype=np.int64)
    pc = _f32(pixel_coords)
    u, v = pc[:, 0].astype(np.int64), pc[:, 1].astype(np.int64)
    dirs = get_ray_directions(H, W, focal)[v, u]
    P = _f32(poses)[img]
    R = P[:, :3, :3]
    prod = dirs[:, None, :] * R
    vw = ((prod[..., 0] + prod[..., 1]) + prod[..., 2]).astype(F32)
    nrm = np.sqrt(((vw[:, 0] ** 2 + vw[:, 1] ** 2) + vw[:, 2] ** 2)).astype(F32)
    rays_d = (vw / nrm[:, None]).astype(F32)
    rays_o = P[:, :3, 3].copy()
    if keep_cache:
        return rays_o, rays_d, dict(img=img, dirs=dirs, nrm=nrm, rays_d=rays_d, n=len(poses))
    return rays_o, rays_d


def get_rays_from_pixels_backward(cache, g_o, g_d):
    """dL/dposes (N,4,4) from dL/drays_o, dL/drays_d."""
    g_o, g_d = _f32(g_o), _f32(g_d)
    d = cache["rays_d"]
    gv = ((g_d - d * (d * g_d).sum(-1, keepdims=True)) / cache["nrm"][:, None]).astype(F32)
    gP = np.zeros((cache["n"], 4, 4), dtype=np.float64)
    np.add.at(gP[:, :3, :3], cache["img"], gv[:, :, None] * cache["dirs"][:, None, :])
    np.add.at(gP[:, :3, 3], cache["img"], g_o)
    return gP.astype(F32)


def compute_pose_error(pose_gt, pose_est):
    """noisy_src/noise.py:237-268: geodesic rotation angle (deg) + translation L2."""
    pose_gt, pose_est = _f32(pose_gt), _f32(pose_est)
    Rdiff = (pose_gt[:3, :3].T @ pose_est[:3, :3]).astype(F32)
    tr = F32((Rdiff[0, 0] + Rdiff[1, 1]) + Rdiff[2, 2])
    ang = np.arccos(np.clip((tr - F32(1.0)) / F32(2.0), F32(-1.0), F32(1.0)))
    dt = pose_gt[:3, 3] - pose_est[:3, 3]
    return {"rotation_error_deg": float(F32(ang) * 180.0 / np.pi),
            "translation_error": float(np.sqrt((dt * dt).sum()))}


def add_noise_to_poses(poses, g_angle, g_axis, g_trans, rotation_noise_deg=0.0, translation_noise=0.0,
                       translation_noise_pct=0.0):
    """noisy_src/noise.py:71-234 (random_rotation_matrix, random_translation, add_noise_to_pose, add_noise_to_poses) given
    the raw standard-normal draws of the reference's generator calls (per pose randn(1), randn(3) for the rotation,
    randn(3) for the translation; None = that noise is off).  fp32 throughout, sums in index order.
    Returns (noisy poses [n,4,4], info [n,2] = actual_rotation_deg, actual_translation_norm)."""
    poses = _f32(poses)
    n = poses.shape[0]
    out, info = poses.copy(), np.zeros((n, 2), F32)
    std_rad = F32(rotation_noise_deg * np.pi / 180.0)
    for i in range(n):
        if g_angle is not None:
            a = F32(F32(g_angle[i]) * std_rad)
            ax = _f32(g_axis[i])
            ax = (ax / np.sqrt(F32(F32(ax[0] * ax[0] + ax[1] * ax[1]) + ax[2] * ax[2]))).astype(F32)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]], F32)
            KK = np.zeros((3, 3), F32)
            for r in range(3):
                for c in range(3):
                    KK[r, c] = F32(F32(K[r, 0] * K[0, c] + K[r, 1] * K[1, c]) + K[r, 2] * K[2, c])
            R = ((np.eye(3, dtype=F32) + np.sin(a).astype(F32) * K).astype(F32) + F32(F32(1.0) - np.cos(a).astype(F32)) * KK).astype(F32)
            Ro = poses[i, :3, :3]
            for r in range(3):
                for c in range(3):
                    out[i, r, c] = F32(F32(R[r, 0] * Ro[0, c] + R[r, 1] * Ro[1, c]) + R[r, 2] * Ro[2, c])
            tr = F32(F32(R[0, 0] + R[1, 1]) + R[2, 2])
            ang = np.arccos(np.clip(F32(tr - F32(1.0)) / F32(2.0), F32(-1.0), F32(1.0))).astype(F32)
            info[i, 0] = F32(F32(ang * F32(180.0)) / F32(np.pi))
        if g_trans is not None:
            t = poses[i, :3, 3]
            dist = np.sqrt(F32(F32(t[0] * t[0] + t[1] * t[1]) + t[2] * t[2])).astype(F32)
            std = F32(float(dist) * (translation_noise_pct / 100.0)) if translation_noise_pct > 0 else F32(translation_noise)
            d = (_f32(g_trans[i]) * std).astype(F32)
            out[i, :3, 3] = (t + d).astype(F32)
            info[i, 1] = np.sqrt(F32(F32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
    return out, info


def compute_pose_errors(poses, gt_poses):
    """noisy_src/train_pose_opt.py:232-271."""
    r, t = zip(*[(e["rotation_error_deg"], e["translation_error"])
                 for e in (compute_pose_error(g, p) for g, p in zip(gt_poses, poses))])
    return {"rotation_error_mean": float(np.mean(r)), "rotation_error_std": float(np.std(r)),
            "rotation_error_max": float(np.max(r)), "translation_error_mean": float(np.mean(t)),
            "translation_error_std": float(np.std(t)), "translation_error_max": float(np.max(t))}


def psnr(pred, target) -> float:
    """noisy_src/metrics.py:15-40: -10 log10(mse)."""
    mse = float(((np.asarray(pred, np.float64) - np.asarray(target, np.float64)) ** 2).mean())
    return float(-10.0 * np.log10(max(mse, 1e-20)))


# --------------------------------------------------------------------------------------
# optimiser tail of the step (train.py:115-117): joint clip at 1.0 + Adam(lr=5e-4)
# --------------------------------------------------------------------------------------
def clip_and_adam(params_list, grads_list, state, lr=5e-4, max_norm=1.0, betas=(0.9, 0.999),
                  eps=1e-8):
    """clip_grad_norm_ over all given tensors jointly, then one torch.optim.Adam step."""
    tot = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for gs in grads_list for g in gs.values()))
    coef = min(1.0, max_norm / (tot + 1e-6))
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    for pi, (ps, gs) in enumerate(zip(params_list, grads_list)):
        for k in ps:
            g = gs[k] * F32(coef)
            m = state.setdefault(("m", pi, k), np.zeros_like(ps[k]))
            v = state.setdefault(("v", pi, k), np.zeros_like(ps[k]))
            m *= F32(betas[0]); m += F32(1 - betas[0]) * g
            v *= F32(betas[1]); v += F32(1 - betas[1]) * g * g
            mhat = m / F32(1 - betas[0] ** t)
            vhat = v / F32(1 - betas[1] ** t)
            ps[k] -= (F32(lr) * mhat / (np.sqrt(vhat) + F32(eps))).astype(F32)
    return tot


# --------------------------------------------------------------------------------------
# evaluation metrics: noisy_src/metrics.py:15-116
# --------------------------------------------------------------------------------------
def gaussian_window_2d(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    """metrics.py:85-89: g = exp(-x^2 / (2 sigma^2)), normalised, outer product (all fp32)."""
    coords = (np.arange(size, dtype=F32) - F32(size // 2)).astype(F32)
    g = np.exp(-(coords ** 2) / F32(2 * sigma ** 2)).astype(F32)
    g = (g / g.sum(dtype=F32)).astype(F32)
    return np.outer(g, g).astype(F32)


def compute_mse(pred, target) -> float:
    """metrics.py:44-46."""
    d = _f32(pred) - _f32(target)
    return float(np.mean((d * d).astype(F32), dtype=np.float64))


def compute_psnr(pred, target, max_val: float = 1.0) -> float:
    """metrics.py:15-41: 20 log10(max) - 10 log10(mse); inf for identical images."""
    mse = compute_mse(pred, target)
    if mse == 0:
        return float("inf")
    return float(20.0 * np.log10(max_val) - 10.0 * np.log10(mse))


def _conv2d_same(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Zero-padded 'same' cross-correlation of x (H, W) with w (k, k), as F.conv2d(padding=k//2) (metrics.py:98-110)."""
    k = w.shape[0]
    r = k // 2
    H, W = x.shape
    xp = np.zeros((H + 2 * r, W + 2 * r), dtype=np.float64)
    xp[r:r + H, r:r + W] = x
    out = np.zeros((H, W), dtype=np.float64)
    for i in range(k):
        for j in range(k):
            out += np.float64(w[i, j]) * xp[i:i + H, j:j + W]
    return out.astype(F32)


def compute_ssim(pred, target, window_size: int = 11, C1: float = 0.01 ** 2, C2: float = 0.03 ** 2) -> float:
    """metrics.py:49-116: per-channel Gaussian-window SSIM map of an (H, W, 3) image pair, mean over all pixels and channels."""
    pred, target = _f32(pred), _f32(target)
    w = gaussian_window_2d(window_size)
    total = 0.0
    for c in range(pred.shape[-1]):
        p, t = pred[..., c], target[..., c]
        mu_p, mu_t = _conv2d_same(p, w), _conv2d_same(t, w)
        mu_pp, mu_tt, mu_pt = mu_p * mu_p, mu_t * mu_t, mu_p * mu_t
        s_pp = _conv2d_same(p * p, w) - mu_pp
        s_tt = _conv2d_same(t * t, w) - mu_tt
        s_pt = _conv2d_same(p * t, w) - mu_pt
        ssim_map = ((2 * mu_pt + F32(C1)) * (2 * s_pt + F32(C2))) / ((mu_pp + mu_tt + F32(C1)) * (s_pp + s_tt + F32(C2)))
        total += float(np.sum(ssim_map, dtype=np.float64))
    return total / pred.size
