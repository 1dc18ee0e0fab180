"""Training / full-image entry points of the hot path.

* `train_step`, `train_step_with_poses`, `render_image`, `render_image_with_pose` keep the
  reference's signatures and metric dictionaries (noisy_src/train.py:68-160,
  noisy_src/train_pose_opt.py:290-470) -- thin callers over the kernel path.
* `Trainer` is the B200-first step: parameters, gradients and Adam moments of both networks live
  in flat fp32 buffers (the nn.Parameters are views, so state_dict / checkpoints are unchanged),
  the batch is sharded data-parallel with ONE NCCL all-reduce per step over the flat gradient
  buffer [grad coarse | grad fine | grad omega | grad delta_t] placed before the clip (so the
  clip sees the global-batch gradient, like a single-GPU run), and clip_grad_norm_ + Adam is
  one fused kernel over the same buffer.  No host synchronisation inside a step.
* `render_views_sharded` shards test-view rendering by ray tile, no collective.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops, _lib as L
from ._lib import call, ptr, stream_ptr
from .config import RenderConfig
from .model import NeRF
from .rays import get_ray_directions, get_rays
from .rendering import NeRFRenderer, render_rays
from .parallel import allreduce_mean_, tiles_for_rank


def compute_psnr(pred: torch.Tensor, target: torch.Tensor, max_val: float = 1.0) -> torch.Tensor:
    """noisy_src/metrics.py:15-40."""
    mse = torch.mean((pred - target) ** 2)
    return 20.0 * torch.log10(torch.tensor(max_val, device=mse.device)) - 10.0 * torch.log10(mse)


class _MSE(torch.autograd.Function):
    """mean((rgb - target)^2) with its gradient produced in the same launch."""

    @staticmethod
    def forward(ctx, rgb_map, target):
        loss = torch.zeros(1, device=rgb_map.device)
        g = ops.mse_loss_and_grad(rgb_map, target, loss, 1.0, want_grad=True)
        ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g * go, None


def mse_loss(rgb_map, target):
    return _MSE.apply(rgb_map, target)


def _losses(outputs, target_rgb):
    loss_coarse = mse_loss(outputs["rgb_coarse"], target_rgb)
    loss_fine = mse_loss(outputs["rgb_fine"], target_rgb) if "rgb_fine" in outputs else None
    return loss_coarse, loss_fine


def train_step(renderer: NeRFRenderer, optimizer: torch.optim.Optimizer, batch: Dict[str, torch.Tensor]) -> Dict[str, float]:
    """Clean-pose step with the reference's semantics (noisy_src/train.py:68-119): joint clip at 1.0."""
    optimizer.zero_grad()
    outputs = renderer(batch["rays_o"], batch["rays_d"], is_train=True)
    target = batch["target_rgb"]
    loss_c, loss_f = _losses(outputs, target)
    loss = loss_c if loss_f is None else loss_c + loss_f
    loss.backward()
    torch.nn.utils.clip_grad_norm_(renderer.parameters(), max_norm=1.0)
    optimizer.step()
    # one host sync for all metrics (the reference does five .item() calls)
    vals = torch.stack([loss_c.detach(), (loss_f if loss_f is not None else loss_c).detach(), loss.detach()]).tolist()
    psnr = lambda m: float("inf") if m == 0 else -10.0 * torch.log10(torch.tensor(m)).item()
    metrics = {"loss_coarse": vals[0], "psnr_coarse": psnr(vals[0])}
    if loss_f is not None:
        metrics.update(loss_fine=vals[1], psnr_fine=psnr(vals[1]), psnr=psnr(vals[1]))
    else:
        metrics.update(loss_fine=None, psnr=metrics["psnr_coarse"])
    metrics["loss"] = vals[2]
    return metrics


def train_step_with_poses(model_coarse: NeRF, model_fine: Optional[NeRF], camera_params, pixel_sampler,
                          optimizer_nerf: torch.optim.Optimizer, optimizer_poses: Optional[torch.optim.Optimizer],
                          pixel_batch, render_config: RenderConfig, optimize_poses: bool = True,
                          rotation_reg_weight: float = 0.0, translation_reg_weight: float = 0.0) -> Dict[str, float]:
    """Joint NeRF + pose step (noisy_src/train_pose_opt.py:290-411): per-net clip 1.0, pose clip 0.1."""
    optimizer_nerf.zero_grad()
    if optimizer_poses is not None and optimize_poses:
        optimizer_poses.zero_grad()
    rays_o, rays_d = pixel_sampler.get_rays_for_batch_fused(pixel_batch, camera_params)
    outputs = render_rays(model_coarse, model_fine, rays_o, rays_d, render_config, is_train=True)
    loss_c, loss_f = _losses(outputs, pixel_batch.target_rgb)
    loss = loss_c if loss_f is None else loss_c + loss_f
    extras = {}
    if optimize_poses and (rotation_reg_weight > 0 or translation_reg_weight > 0):
        reg = 0.0
        if rotation_reg_weight > 0 and camera_params.learn_rotation:
            r = torch.mean(camera_params.rotation_deltas ** 2)
            reg = reg + rotation_reg_weight * r
            extras["rotation_reg"] = r.detach()
        if translation_reg_weight > 0 and camera_params.learn_translation:
            t = torch.mean(camera_params.translation_deltas ** 2)
            reg = reg + translation_reg_weight * t
            extras["translation_reg"] = t.detach()
        loss = loss + reg
        extras["pose_reg_loss"] = reg.detach() if isinstance(reg, torch.Tensor) else torch.tensor(float(reg))
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model_coarse.parameters(), max_norm=1.0)
    if model_fine is not None:
        torch.nn.utils.clip_grad_norm_(model_fine.parameters(), max_norm=1.0)
    if optimize_poses and optimizer_poses is not None:
        torch.nn.utils.clip_grad_norm_(camera_params.parameters(), max_norm=0.1)
    optimizer_nerf.step()
    if optimizer_poses is not None and optimize_poses:
        optimizer_poses.step()
    names = ["loss_coarse", "loss_fine", "loss"] + list(extras)
    vals = torch.stack([loss_c.detach(), (loss_f if loss_f is not None else loss_c).detach(), loss.detach()]
                       + [v.to(loss.device) for v in extras.values()]).tolist()
    m = dict(zip(names, vals))
    psnr = lambda x: float("inf") if x == 0 else -10.0 * torch.log10(torch.tensor(x)).item()
    m["psnr_coarse"] = psnr(m["loss_coarse"])
    if loss_f is not None:
        m["psnr_fine"] = psnr(m["loss_fine"])
        m["psnr"] = m["psnr_fine"]
    else:
        m["loss_fine"] = None
        m["psnr"] = m["psnr_coarse"]
    return m


_U_DET_CACHE = {}


def _eval_rows(cfg: RenderConfig, device):
    """The two linspace rows of an evaluation render (rays.py:185-192 base depths, rays.py:252 deterministic draws),
    built with the reference's torch calls on this device, once."""
    from .rays import _z_base
    zb = _z_base(cfg.near, cfg.far, cfg.num_samples, False, device)
    key = (int(cfg.num_samples_fine), str(device))
    u = _U_DET_CACHE.get(key)
    if u is None:
        u = torch.linspace(0.0, 1.0, cfg.num_samples_fine, device=device)
        if not torch.cuda.is_current_stream_capturing():
            _U_DET_CACHE[key] = u
    return zb, u


def _fused_render_ok(model_coarse, model_fine) -> bool:
    return isinstance(model_coarse, NeRF) and (model_fine is None or isinstance(model_fine, NeRF))


@torch.no_grad()
def render_image(renderer: NeRFRenderer, pose: torch.Tensor, H: int, W: int, focal: float,
                 chunk_size: int = 1024 * 4) -> Dict[str, torch.Tensor]:
    """noisy_src/train.py:122-160.  One C-ABI call (`rn_render_view`): the rays are generated from the pose inside it and
    the image is rendered in tiles without a Python loop; `chunk_size` is accepted for signature parity (results are
    chunk-invariant in eval mode; the tile is max(chunk_size, 131072) rays)."""
    cfg = renderer.config
    mc, mf = renderer.model_coarse, (renderer.model_fine if cfg.use_hierarchical else None)
    if _fused_render_ok(mc, mf):
        zb, u = _eval_rows(cfg, pose.device)
        rgb, depth, acc, _ = ops.render_view(mc, mf, pose, H, W, focal, zb, u if mf is not None else None,
                                             white_background=cfg.white_background, tile_rays=max(int(chunk_size), 131072))
        return {"rgb": rgb.reshape(H, W, 3), "depth": depth.reshape(H, W), "acc": acc.reshape(H, W)}
    directions = get_ray_directions(H, W, focal, device=pose.device)
    rays_o, rays_d = get_rays(directions, pose)
    out = renderer(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), chunk_size=chunk_size, is_train=False)
    k = "fine" if "rgb_fine" in out else "coarse"
    return {"rgb": out[f"rgb_{k}"].reshape(H, W, 3), "depth": out[f"depth_{k}"].reshape(H, W),
            "acc": out[f"acc_{k}"].reshape(H, W)}


@torch.no_grad()
def render_image_with_pose(model_coarse: NeRF, model_fine: Optional[NeRF], pose: torch.Tensor, H: int, W: int,
                           focal: float, render_config: RenderConfig, chunk_size: int = 1024 * 4) -> Dict[str, torch.Tensor]:
    """noisy_src/train_pose_opt.py:414-470."""
    return render_image(NeRFRenderer(model_coarse, model_fine, render_config), pose, H, W, focal, chunk_size)


# ------------------------------------------------------------------------------------------------
# B200-first trainer
# ------------------------------------------------------------------------------------------------
def _flatten_into(params: Sequence[nn.Parameter], flat: torch.Tensor, gflat: torch.Tensor, off: int) -> int:
    """Re-point every parameter (and its .grad) at a slice of the flat buffers; values preserved."""
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        p.grad = gflat[off:off + n].view(p.shape)
        off += n
    return off


class Trainer:
    """Fused render+train step with optional data parallelism.

    clean mode  : step_rays(rays_o, rays_d, target)                 (train.py:68-119 semantics, joint clip 1.0)
    pose mode   : step_pixels(pixel_batch, sampler, optimize_poses) (train_pose_opt.py:290-411 semantics:
                  per-net clip 1.0, pose clip 0.1, separate Adam at pose_lr, optional L2 pose regulariser)
    Under torch.distributed (one process per GPU) each rank passes its own shard of the batch; gradients are
    averaged with one all-reduce over the flat buffer.
    """

    def __init__(self, model_coarse: NeRF, model_fine: Optional[NeRF], render_config: RenderConfig, lr: float = 5e-4,
                 lr_decay_steps: float = 250000.0, camera_params=None, pose_lr: float = 1e-4,
                 rotation_reg_weight: float = 0.0, translation_reg_weight: float = 0.0, process_group=None,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.model_coarse, self.model_fine, self.cfg = model_coarse, model_fine, render_config
        self.camera_params = camera_params
        self.lr, self.pose_lr, self.lr_decay_steps = lr, pose_lr, lr_decay_steps
        self.rot_reg, self.trans_reg = rotation_reg_weight, translation_reg_weight
        self.betas, self.eps = betas, eps
        self.group = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        nets = [model_coarse] + ([model_fine] if model_fine is not None else [])
        dev = next(model_coarse.parameters()).device
        self.device = dev
        # Flat layout [fine | coarse | omega | delta_t]: autograd runs the fine network's backward first, so its slice is
        # final while the coarse backward still runs and can be all-reduced underneath it; what is left afterwards
        # (coarse + poses) is one contiguous range.  `net_params` keeps the reference's optimiser order (coarse, fine).
        flat_order = list(reversed(nets))
        self.net_params: List[List[nn.Parameter]] = [m.kernel_params() for m in nets]
        self.pose_params: List[nn.Parameter] = list(camera_params.parameters()) if camera_params is not None else []
        n_net = sum(p.numel() for ps in self.net_params for p in ps)
        n_pose = sum(p.numel() for p in self.pose_params)
        self.n_net, self.n_pose = n_net, n_pose
        total = n_net + n_pose
        self.flat = torch.zeros(total, device=dev)
        self.gflat = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        off = 0
        self.net_offsets = [0]                       # boundaries of the nets in FLAT order (clip groups)
        self._sinks = {}                             # net -> its slice of gflat
        self._param_off = {}                         # id(param) -> offset in the flat buffers
        for m in flat_order:
            start = off
            for q in m.kernel_params():
                self._param_off[id(q)] = off + 0
                off += q.numel()
            _flatten_into(m.kernel_params(), self.flat, self.gflat, start)
            self.net_offsets.append(off)
            self._sinks[m] = self.gflat[start:off]   # armed as the net's gradient sink only inside a Trainer step
        self._early = (0, self.net_offsets[1]) if (len(nets) == 2 and self.world > 1) else None   # fine slice
        for q in self.pose_params:
            self._param_off[id(q)] = off
            off += q.numel()
        _flatten_into(self.pose_params, self.flat, self.gflat, n_net)
        self._side = torch.cuda.Stream(device=dev) if self._early else None
        self._early_issued = False
        self.skip_allreduce = False        # measurement only (bench.py: exposed time of the collectives); replicas diverge while set
        self.norms = torch.zeros(2 * (8 + 8 * 64), device=dev)    # two rn_clip_adam_step scratch areas
        # [lr, 1-b1^t, sqrt(1-b2^t)] for the nets and for the poses: read by the Adam kernel from device memory so
        # a captured CUDA graph can be replayed while the schedule advances
        # Pinned ring of schedule rows: the async H2D copy of step t reads slot t % RING; the slot is rewritten at step
        # t + RING, so each slot carries an event recorded after its copy and the host waits on it before reuse (a host
        # that runs more than RING replays ahead of the device would otherwise corrupt an in-flight step's LR).
        self.hyper_host = torch.zeros(self.RING, 6).pin_memory()
        self._hyper_events: List[Optional[torch.cuda.Event]] = [None] * self.RING
        self.hyper = torch.zeros(6, device=dev)
        self._graphs = {}
        self.iteration = 0
        self.pose_steps = 0
        self.nets = nets

    RING = 512

    # -- pieces --------------------------------------------------------------------------------
    def _arm_sinks(self, on: bool):
        """Point the nets' backward kernels at the flat gradient buffer for the duration of a Trainer step only:
        outside it the models behave like ordinary modules (p.grad filled by autograd), so the reference-style
        train_step / train_step_with_poses on the same models keep working."""
        for m in self.nets:
            m._packed.grad_sink = self._sinks[m] if on else None
            m._packed.sink_dirty = False
            m._packed.on_grads_ready = None
        if on:
            self._early_issued = False           # (cleared again by _allreduce, which runs after the sinks are disarmed)
            if self._early is not None:
                self.model_fine._packed.on_grads_ready = self._allreduce_early

    def _allreduce_early(self):
        """Called by the fine network's backward as soon as its gradient slice is final: all-reduce it on a side stream,
        underneath the coarse network's backward (captured into the step's CUDA graph like everything else)."""
        if self.skip_allreduce:
            return
        lo, hi = self._early
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            torch.distributed.all_reduce(self.gflat[lo:hi], op=torch.distributed.ReduceOp.SUM, group=self.group)
        self._early_issued = True

    def _neutral_hyper(self):
        """lr = 0 with unit bias corrections: warm-up / capture passes leave the parameters untouched (all-zero rows
        would divide 0 by 0 in the Adam kernel and push NaN weights through the whole pipeline)."""
        self.hyper.copy_(torch.tensor([0.0, 1.0, 1.0, 0.0, 1.0, 1.0], device=self.hyper.device))

    def _zero_grad(self):
        self._check_aliasing()
        if self.n_pose:
            self.gflat[self.n_net:].zero_()                  # net gradients are overwritten by the kernels
        for p in [q for ps in self.net_params for q in ps] + self.pose_params:
            off, n = self._param_off[id(p)], p.numel()
            if p.grad is None or p.grad.data_ptr() != self.gflat.data_ptr() + 4 * off:
                p.grad = self.gflat[off:off + n].view(p.shape)   # somebody called zero_grad(set_to_none=True)

    def _allreduce(self):
        """SUM over ranks of whatever has not been reduced yet; the 1 / world of the mean is applied by the clip + Adam
        kernel on load (`grad_scale`), not by another launch."""
        if self.world <= 1 or self.skip_allreduce:
            return
        if self._early_issued:
            lo = self._early[1]
            torch.distributed.all_reduce(self.gflat[lo:], op=torch.distributed.ReduceOp.SUM, group=self.group)
            torch.cuda.current_stream().wait_stream(self._side)
            self._early_issued = False
        else:
            torch.distributed.all_reduce(self.gflat, op=torch.distributed.ReduceOp.SUM, group=self.group)

    def _adam(self, lo: int, hi: int, groups: Sequence[int], max_norms: Sequence[float], hyper_slot: int, norm_slot: int):
        n = hi - lo
        offs = (ctypes.c_int64 * len(groups))(*[g - lo for g in groups])
        mx = (ctypes.c_float * len(max_norms))(*max_norms)
        call("rn_clip_adam_step", self.flat[lo:hi].data_ptr(), self.gflat[lo:hi].data_ptr(), self.exp_avg[lo:hi].data_ptr(),
             self.exp_avg_sq[lo:hi].data_ptr(), n, offs, mx, len(max_norms), 0.0, float(self.betas[0]),
             float(self.betas[1]), float(self.eps), 1, self.norms[norm_slot:].data_ptr(),
             self.hyper[hyper_slot:].data_ptr(), 1.0 / self.world, stream_ptr())

    def _advance_schedule(self, optimize_poses: bool):
        """Host side of the optimiser schedule (train.py:405-411: lr * 0.1^(step/250k)); uploaded to the device
        buffer the Adam kernel reads, outside any captured graph."""
        self.iteration += 1
        t = self.iteration
        ev = self._hyper_events[t % self.RING]
        if ev is not None:
            ev.synchronize()                                  # the copy issued RING steps ago has read this slot
        h = self.hyper_host[t % self.RING]
        h[3:] = self.hyper_host[(t - 1) % self.RING][3:]
        if t == 1:
            h[3:] = torch.tensor([0.0, 1.0, 1.0])
        h[0] = self.lr * (0.1 ** ((t - 1) / self.lr_decay_steps))
        h[1] = 1.0 - self.betas[0] ** t
        h[2] = (1.0 - self.betas[1] ** t) ** 0.5
        if self.n_pose and optimize_poses:
            self.pose_steps += 1
            tp = self.pose_steps
            h[3] = self.pose_lr * (0.1 ** ((tp - 1) / self.lr_decay_steps))
            h[4] = 1.0 - self.betas[0] ** tp
            h[5] = (1.0 - self.betas[1] ** tp) ** 0.5
        self.hyper.copy_(h, non_blocking=True)
        ev = self._hyper_events[t % self.RING] or torch.cuda.Event()
        ev.record()
        self._hyper_events[t % self.RING] = ev

    def _optimise(self, separate_clip: bool, optimize_poses: bool):
        if separate_clip and len(self.nets) == 2:
            self._adam(0, self.n_net, self.net_offsets, [1.0, 1.0], 0, 0)
        else:
            self._adam(0, self.n_net, [0, self.n_net], [1.0], 0, 0)
        if self.n_pose and optimize_poses:
            self._adam(self.n_net, self.n_net + self.n_pose, [self.n_net, self.n_net + self.n_pose], [0.1], 3, 8 + 8 * 64)
        for m in self.nets:
            m._packed.key = None          # parameters changed under the bf16 cache

    def _backward(self, outputs, target, extra_loss=None):
        """Loss and the start of the backward pass without an autograd node for the loss: one kernel produces
        mse(coarse) + mse(fine) and both gradients w.r.t. the rendered colours (train.py:88-99), which are fed straight
        into autograd.backward on the two rgb maps (the optional pose regulariser, a scalar, rides along)."""
        rgb_c, rgb_f = outputs["rgb_coarse"], outputs.get("rgb_fine")
        loss3, g_c, g_f = ops.mse2_loss_and_grads(rgb_c, rgb_f, target)
        tensors, grads = [rgb_c], [g_c]
        if rgb_f is not None:
            tensors.insert(0, rgb_f)             # fine first: its gradient slice is then final first (early all-reduce)
            grads.insert(0, g_f)
        loss = loss3[0]
        if extra_loss is not None:
            tensors.append(extra_loss)
            grads.append(torch.ones_like(extra_loss))
            loss = loss + extra_loss.detach()
        torch.autograd.backward(tensors, grads)
        return loss

    # -- checkpointing (train.py:248-271, train_pose_opt.py:563-597: optimizer.state_dict() travels with the model) ------
    def _adam_state(self, params: Sequence[nn.Parameter], base_lr: float, steps: int) -> dict:
        state = {}
        if steps > 0:
            for i, p in enumerate(params):
                off, n = self._param_off[id(p)], p.numel()
                state[i] = {"step": torch.tensor(float(steps)),
                            "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        group = {"lr": base_lr * (0.1 ** (steps / self.lr_decay_steps)), "betas": tuple(self.betas), "eps": self.eps,
                 "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": False, "initial_lr": base_lr,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def state_dict(self) -> dict:
        """Optimiser state in `torch.optim.Adam.state_dict()` format (per-parameter `step`, `exp_avg`, `exp_avg_sq`, in
        the reference's parameter order: coarse net then fine net; rotation then translation deltas), so a reference
        checkpoint's `optimizer_state_dict` loads here and ours loads into `torch.optim.Adam`.  The models and the camera
        parameters are saved through their own `state_dict()` as in the reference."""
        net_params = [p for ps in self.net_params for p in ps]
        out = {"optimizer_nerf": self._adam_state(net_params, self.lr, self.iteration), "iteration": self.iteration,
               "pose_steps": self.pose_steps}
        if self.n_pose:
            out["optimizer_poses"] = self._adam_state(self.pose_params, self.pose_lr, self.pose_steps)
        return out

    def _load_adam_state(self, sd: dict, params: Sequence[nn.Parameter]) -> int:
        steps = 0
        for i, p in enumerate(params):
            off, n = self._param_off[id(p)], p.numel()
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                self.exp_avg[off:off + n].zero_(); self.exp_avg_sq[off:off + n].zero_()
            else:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {i} has shape {tuple(st['exp_avg'].shape)}, parameter has {tuple(p.shape)}")
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps = max(steps, int(float(st["step"])))
        return steps

    def load_state_dict(self, sd: dict) -> None:
        """Accepts `Trainer.state_dict()` or a dict holding `torch.optim.Adam.state_dict()`s under the same keys."""
        net_params = [p for ps in self.net_params for p in ps]
        steps = self._load_adam_state(sd["optimizer_nerf"], net_params)
        self.iteration = int(sd.get("iteration", steps))
        if self.n_pose and "optimizer_poses" in sd:
            psteps = self._load_adam_state(sd["optimizer_poses"], self.pose_params)
            self.pose_steps = int(sd.get("pose_steps", psteps))
        # schedule rows are rebuilt from the step counters on the next _advance_schedule; the pose row must survive a
        # clean-mode step, so seed the previous slot with it
        h = self.hyper_host[self.iteration % self.RING]
        tp = max(self.pose_steps, 1)
        h[3] = self.pose_lr * (0.1 ** ((tp - 1) / self.lr_decay_steps))
        h[4] = 1.0 - self.betas[0] ** tp
        h[5] = (1.0 - self.betas[1] ** tp) ** 0.5
        self._check_aliasing()

    def _check_aliasing(self):
        """Parameters must still be views of the flat buffer (model.to(), p.data = ... after construction detaches them)."""
        for p in [q for ps in self.net_params for q in ps] + self.pose_params:
            if p.data_ptr() != self.flat.data_ptr() + 4 * self._param_off[id(p)]:
                raise RuntimeError("a parameter no longer aliases the Trainer's flat buffer (was the model moved or "
                                   "re-assigned after the Trainer was built?); rebuild the Trainer")

    # -- public steps -----------------------------------------------------------------------------
    def _step_rays_body(self, rays_o, rays_d, target, optimise):
        self._zero_grad()
        self._arm_sinks(True)
        try:
            out = render_rays(self.model_coarse, self.model_fine, rays_o, rays_d, self.cfg, is_train=True)
            loss = self._backward(out, target)
        finally:
            self._arm_sinks(False)
        self._allreduce()
        if optimise:
            self._optimise(separate_clip=False, optimize_poses=False)
        return loss

    def step_rays(self, rays_o: torch.Tensor, rays_d: torch.Tensor, target: torch.Tensor, optimise: bool = True
                  ) -> torch.Tensor:
        if optimise:
            self._advance_schedule(False)
        return self._step_rays_body(rays_o, rays_d, target, optimise)

    def step_rays_graphed(self, rays_o: torch.Tensor, rays_d: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """step_rays replayed from a CUDA graph (captured on first use for this batch size; single GPU or DP --
        the NCCL all-reduce is captured with the rest).  Inputs are copied into static buffers; the ~110 kernel
        launches of a step then cost one graph launch.  Returns a static loss tensor (overwritten by the next step)."""
        key = ("rays", tuple(rays_o.shape))
        g = self._graphs.get(key)
        if g is None:
            static = [torch.empty_like(rays_o), torch.empty_like(rays_d), torch.empty_like(target)]
            for s_, x in zip(static, (rays_o, rays_d, target)):
                s_.copy_(x)
            # keep parameters / optimiser state untouched by warm-up and capture
            saved = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq)]
            self._neutral_hyper()                    # lr = 0, unit bias corrections during warm-up/capture passes
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step_rays_body(*static, True)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = self._step_rays_body(*static, True)
            for t, sv in zip((self.flat, self.exp_avg, self.exp_avg_sq), saved):
                t.copy_(sv)
            for m in self.nets:
                m._packed.key = None
            g = self._graphs[key] = (graph, static, loss)
        graph, static, loss = g
        for s_, x in zip(static, (rays_o, rays_d, target)):
            s_.copy_(x, non_blocking=True)
        self._advance_schedule(False)
        graph.replay()
        return loss

    def _step_pixels_body(self, pixel_batch, sampler, optimize_poses: bool, optimise: bool, advance: bool):
        cam = self.camera_params
        self._zero_grad()
        self._arm_sinks(True)
        try:
            rays_o, rays_d = sampler.get_rays_for_batch_fused(pixel_batch, cam)
            out = render_rays(self.model_coarse, self.model_fine, rays_o, rays_d, self.cfg, is_train=True)
            reg = None
            if optimize_poses and (self.rot_reg > 0 or self.trans_reg > 0):
                reg = 0.0
                if self.rot_reg > 0 and cam.learn_rotation:
                    reg = reg + self.rot_reg * torch.mean(cam.rotation_deltas ** 2)
                if self.trans_reg > 0 and cam.learn_translation:
                    reg = reg + self.trans_reg * torch.mean(cam.translation_deltas ** 2)
            loss = self._backward(out, pixel_batch.target_rgb, reg)
        finally:
            self._arm_sinks(False)
        self._allreduce()
        if optimise:
            if advance:
                self._advance_schedule(optimize_poses)
            self._optimise(separate_clip=True, optimize_poses=optimize_poses)
        return loss

    def step_pixels(self, pixel_batch, sampler, optimize_poses: bool = True, optimise: bool = True) -> torch.Tensor:
        return self._step_pixels_body(pixel_batch, sampler, optimize_poses, optimise, advance=True)

    def step_pixels_graphed(self, pixel_batch, sampler, optimize_poses: bool = True) -> torch.Tensor:
        """step_pixels (joint pose optimisation, train_pose_opt.py:290-411) replayed from a CUDA graph, captured on first
        use per (batch size, optimize_poses).  Same contract as step_rays_graphed: inputs are copied into static
        buffers, the optimiser schedule is uploaded outside the graph, the returned loss tensor is static."""
        from .data_pose_opt import PixelBatch
        key = ("pixels", int(pixel_batch.image_indices.shape[0]), bool(optimize_poses))
        g = self._graphs.get(key)
        if g is None:
            static = PixelBatch(image_indices=pixel_batch.image_indices.clone(), pixel_coords=pixel_batch.pixel_coords.clone(),
                                target_rgb=pixel_batch.target_rgb.clone())
            saved = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq)]
            self._neutral_hyper()                    # lr = 0, unit bias corrections during warm-up/capture passes
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step_pixels_body(static, sampler, optimize_poses, True, advance=False)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = self._step_pixels_body(static, sampler, optimize_poses, True, advance=False)
            for t, sv in zip((self.flat, self.exp_avg, self.exp_avg_sq), saved):
                t.copy_(sv)
            for m in self.nets:
                m._packed.key = None
            g = self._graphs[key] = (graph, static, loss)
        graph, static, loss = g
        static.image_indices.copy_(pixel_batch.image_indices, non_blocking=True)
        static.pixel_coords.copy_(pixel_batch.pixel_coords, non_blocking=True)
        static.target_rgb.copy_(pixel_batch.target_rgb, non_blocking=True)
        self._advance_schedule(optimize_poses)
        graph.replay()
        return loss


@torch.no_grad()
def evaluate(renderer: NeRFRenderer, val_data, num_images: int = 5, host_images: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """The numeric core of noisy_src/train.py:163-233 (`evaluate`; the logger and LPIPS stay with the caller): the first
    `num_images` validation views rendered (one C-ABI call each), MSE / PSNR / SSIM of ALL of them in one metrics launch,
    ONE device-to-host copy of the per-image values (the reference does three `.item()` round trips per image).
    `host_images` (optional pinned (n, H, W, 3) tensor): the renders are copied into it asynchronously, each view's copy
    overlapping the next view's render.  Keys follow ValidationMetrics: psnr, ssim, mse, per_image_psnr, per_image_ssim."""
    from .metrics import image_metrics
    n = min(num_images, val_data.images.shape[0])
    H, W = int(val_data.H), int(val_data.W)
    dev = val_data.poses.device
    preds = torch.empty(n, H, W, 3, device=dev)
    depths = []
    copy_stream = torch.cuda.Stream(device=dev) if host_images is not None else None
    for i in range(n):
        out = render_image(renderer, val_data.poses[i], H, W, float(val_data.focal))
        preds[i] = out["rgb"]
        depths.append(out["depth"])
        if host_images is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                host_images[i].copy_(preds[i], non_blocking=True)
    target = val_data.images[:n]
    if target.dtype == torch.uint8:
        target = ops.dequantize_images(target)
    m = image_metrics(preds, target.to(dev, torch.float32))
    vals = torch.stack([m["mse"], m["psnr"], m["ssim"]]).cpu()             # the one synchronising copy
    if copy_stream is not None:
        copy_stream.synchronize()
    mse, psnr, ssim = vals[0].tolist(), vals[1].tolist(), vals[2].tolist()
    mean = lambda x: float(sum(x) / max(len(x), 1))
    return {"psnr": mean(psnr), "ssim": mean(ssim), "mse": mean(mse), "lpips": None, "per_image_psnr": psnr,
            "per_image_ssim": ssim, "pred": preds, "depth": depths}


@torch.no_grad()
def render_views_sharded(model_coarse: NeRF, model_fine: Optional[NeRF], poses: torch.Tensor, H: int, W: int, focal: float,
                         render_config: RenderConfig, tile_rays: int = 32768, rank: int = 0, world: int = 1,
                         out: Optional[torch.Tensor] = None, view_offset: int = 0) -> Dict[str, torch.Tensor]:
    """Test-view rendering sharded by ray tile: global tile t = view * tiles_per_view + k goes to rank
    t % world; no collective.  Returns this rank's tiles written into a (n_views, H*W, 3) buffer
    (zeros elsewhere) and the number of rays this rank rendered.  `view_offset`: index of poses[0] in the whole job
    (callers that render a long test set a few views at a time keep the global round-robin)."""
    n_views = poses.shape[0]
    npix = H * W
    tiles_per_view = (npix + tile_rays - 1) // tile_rays
    if out is None:
        out = torch.zeros(n_views, npix, 3, device=poses.device)
    mf = model_fine if render_config.use_hierarchical else None
    done = 0
    if _fused_render_ok(model_coarse, mf):
        # one C-ABI call per view: rays from the pose inside the call, this rank's tiles only (global tile id =
        # (view + view_offset) * tiles_per_view + k, dealt round-robin: the first owned tile of a view and the stride follow)
        zb, u = _eval_rows(render_config, poses.device)
        poses = poses.contiguous()
        for v in range(n_views):
            first = (rank - (v + view_offset) * tiles_per_view) % world
            if first >= tiles_per_view:
                continue
            _, _, _, n = ops.render_view(model_coarse, mf, poses[v], H, W, focal, zb, u if mf is not None else None,
                                         white_background=render_config.white_background, tile_rays=tile_rays,
                                         tile_first=first, tile_step=world, out=out[v], want_depth_acc=False)
            done += n
        return {"rgb": out, "rays_rendered": done}
    directions = get_ray_directions(H, W, focal, device=poses.device).reshape(-1, 3)
    for v in range(n_views):
        mine = tiles_for_rank(v + view_offset, tiles_per_view, rank, world)
        if not mine:
            continue
        for k in mine:
            a, b = k * tile_rays, min((k + 1) * tile_rays, npix)
            ro, rd = get_rays(directions[a:b], poses[v])
            res = render_rays(model_coarse, model_fine, ro, rd, render_config, is_train=False)
            out[v, a:b] = res["rgb_fine"] if "rgb_fine" in res else res["rgb_coarse"]
            done += b - a
    return {"rgb": out, "rays_rendered": done}
