"""GPU parity at the benchmark's own size, against the fp32 PyTorch restatement (oracle/torch_ref.py, pinned to the
reference's golden vectors in tests/test_torch_ref_golden.py) running on the same B200:

(a) inverse-CDF sample indices of the CUDA kernel against the reference's own `inds` on 4096 rays (64/128) and
    2048 rays (128/256): mismatch rate <= 2e-5, printed;
(b) per-tensor relative L2 and cosine of every parameter gradient of both networks, and of the pose gradients,
    at 4096 rays (1 M points), on random-init weights and on weights after 200 training steps;
(c) 300-step convergence on a learnable synthetic scene (targets rendered from a fixed teacher network): PSNR
    trajectory against the fp32 restatement trained on the same batches and random draws; pose-error trajectory in
    joint pose-optimisation mode with omega seeded at 1e-3 (live rotation branch, quirk 11).

Every measured number is printed (pytest -s) and written to gpurun_out/parity_r02.json.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden, GOLDEN, ROOT
from oracle import nerf_oracle as O
from oracle import torch_ref as TR

sys.path.insert(0, GOLDEN)
from make_golden_pdf_big import CASES, pdf_big_inputs  # noqa: E402

pytestmark = pytest.mark.gpu

_REPORT = {}


def _report(key, value):
    _REPORT[key] = value
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "parity_r02.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old[key] = value
        json.dump(old, open(path, "w"), indent=1)
    except Exception:
        pass


@pytest.fixture(scope="module")
def rn():
    import robust_nerf_b200 as m
    return m


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _net_from(rn, weights, dev):
    net = rn.NeRF().to(dev)
    sd = net.state_dict()
    for k, v in weights.items():
        sd[k] = torch.as_tensor(v).to(dev)
    net.load_state_dict(sd)
    return net


def _weights_of(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items() if "freq_bands" not in k}


# ------------------------------------------------------------------------------------------------
# (a) index contract against the reference itself
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", list(CASES))
def test_sample_pdf_indices_vs_reference_at_benchmark_size(rn, dev, case):
    from robust_nerf_b200 import ops
    B, Nc, Nf, _ = CASES[case]
    g = load_golden("sample_pdf_big")
    z, w, u = pdf_big_inputs(case)
    inds_ref = g[f"{case}_inds"].astype(np.int64)
    mids = (np.float32(0.5) * (z[..., 1:] + z[..., :-1])).astype(np.float32)
    s, i = ops.sample_pdf(T(mids, dev), T(w[..., 1:-1].copy(), dev), T(u, dev), return_inds=True)
    i = i.cpu().numpy()
    s_o, i_o, c_o = O.sample_pdf(mids, w[..., 1:-1], Nf, det=False, u=u, return_inds=True)
    assert np.array_equal(i, i_o) and np.array_equal(s.cpu().numpy(), s_o)      # kernel == oracle, bit for bit
    mism = i != inds_ref
    rate = float(mism.mean())
    print(f"[{case}] CUDA kernel vs the reference's inds: {int(mism.sum())}/{mism.size} mismatches = {rate:.2e}")
    _report(f"sample_pdf_{case}", {"draws": int(mism.size), "index_mismatches": int(mism.sum()), "rate": rate})
    assert rate <= 2e-5
    if mism.any():
        assert int(np.abs(i - inds_ref)[mism].max()) == 1
    # the hierarchical entry point (merge included) against the torch restatement on this device
    ro = np.zeros((B, 3), np.float32); rd = np.tile(np.array([[0.0, 0.6, -0.8]], np.float32), (B, 1))
    z_all, _, _ = ops.sample_hierarchical(T(ro, dev), T(rd, dev), T(z, dev), T(w, dev), T(u, dev), want_pts=False)
    _, z_t = TR.sample_hierarchical(T(ro, dev), T(rd, dev), T(z, dev), T(w, dev), Nf, u=T(u, dev))
    bad = (z_all - z_t).abs() > 2e-5 + 1e-5 * z_t.abs()
    print(f"[{case}] merged z vs torch restatement on the GPU: {int(bad.sum())}/{bad.numel()} beyond 1e-5 rel")
    assert float(bad.float().mean()) < 1e-3


# ------------------------------------------------------------------------------------------------
# (b) gradients at 4096 rays
# ------------------------------------------------------------------------------------------------
def _teacher_targets(pc, pf, ro, rd, chunk=8192):
    out = []
    with torch.no_grad():
        for a in range(0, ro.shape[0], chunk):
            out.append(TR.render_rays(pc, pf, ro[a:a + chunk], rd[a:a + chunk], is_train=False)["rgb_fine"])
    return torch.cat(out, 0)


def _scene_rays(rn, dev, n, seed):
    """n random rays over the 100 lego cameras (800x800 intrinsics)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    H = W = 800
    poses = rn.lego_poses(dev)
    idx = torch.randint(0, 100 * H * W, (n,), generator=g).to(dev)
    img, uv, _ = rn.ops.pixel_gather(idx, H, W, None)
    dirs = TR.get_ray_directions(H, W, TR.lego_focal(W), device=dev)
    ro, rd = TR.rays_from_pixels(img, uv, poses, dirs)
    return ro.contiguous(), rd.contiguous(), img, uv


def _cmp(a: torch.Tensor, b: torch.Tensor):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    rel = float((a - b).norm() / b.norm().clamp_min(1e-300))
    cos = float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
    return rel, cos


def _grad_table(rn, dev, wc, wf, B, seed, label):
    """Our kernels vs the fp32 restatement: same weights, rays, targets and random draws."""
    ro, rd, _, _ = _scene_rays(rn, dev, B, seed)
    g = torch.Generator(device=dev).manual_seed(seed)
    target = torch.rand(B, 3, device=dev, generator=g)
    t_rand = torch.rand(B, 64, device=dev, generator=g)
    u = torch.rand(B, 128, device=dev, generator=g)
    nc, nf = _net_from(rn, wc, dev), _net_from(rn, wf, dev)
    out = rn.render_rays(nc, nf, ro, rd, rn.RenderConfig(), is_train=True, t_rand=t_rand, u=u, return_extras=True)
    loss = ((out["rgb_coarse"] - target) ** 2).mean() + ((out["rgb_fine"] - target) ** 2).mean()
    loss.backward()
    pc, pf = TR.to_params(wc, dev), TR.to_params(wf, dev)
    res = TR.render_rays(pc, pf, ro, rd, is_train=True, t_rand=t_rand, u=u)
    loss_ref = TR.render_loss(res, target)
    loss_ref.backward()
    rows = {}
    for tag, net, p in (("coarse", nc, pc), ("fine", nf, pf)):
        for k, v in net.named_parameters():
            rows[f"{tag}.{k}"] = _cmp(v.grad, p[k].grad)
        flat_a = torch.cat([v.grad.reshape(-1) for _, v in net.named_parameters()])
        flat_b = torch.cat([p[k].grad.reshape(-1) for k, _ in net.named_parameters()])
        rows[f"{tag}.ALL"] = _cmp(flat_a, flat_b)
    # RGB parity.  The reference's compositor gives the LAST sample an interval of 1e10 (rendering.py:69-72), so its weight is
    # T_last * (sigma_last > 0): a step function of the last density.  A ray whose last pre-activation sits within bf16
    # rounding of zero can therefore differ by up to T_last in ANY reduced-precision forward (it flips in an fp32 forward
    # with another summation order too).  Such rays are counted and must be explained by exactly that: everything but
    # the last sample's weight agrees.
    diff = (out["rgb_fine"].detach() - res["rgb_fine"].detach()).abs().max(-1)[0]
    flipped = diff > 1e-2
    n_flip = int(flipped.sum())
    if n_flip:
        wo, wr = out["weights_fine"].detach()[flipped], res["weights_fine"].detach()[flipped]
        assert float((wo[:, :-1] - wr[:, :-1]).abs().max()) < 1e-2, "rgb outlier not explained by the last sample's density sign"
        assert float((wo[:, -1] - wr[:, -1]).abs().min()) > 1e-2
    rgb_err = float(diff[~flipped].max())
    print(f"\n[{label}] {B} rays: loss {float(loss):.6f} (fp32 restatement {float(loss_ref):.6f}), rgb_fine max-abs {rgb_err:.2e} "
          f"over {B - n_flip} rays; {n_flip} ray(s) flipped by the sign of the last sample's density (1e10 interval)")
    print(f"{'tensor':38s} {'rel L2':>9s} {'cosine':>9s}")
    for k, (rel, cos) in rows.items():
        print(f"{k:38s} {rel:9.4f} {cos:9.5f}")
    _report(f"grad_parity_{label}", {"rays": B, "loss": float(loss), "loss_fp32": float(loss_ref), "rgb_fine_max_abs": rgb_err,
                                    "rays_flipped_by_last_sample_density_sign": n_flip,
                                    "per_tensor": {k: {"rel_l2": r, "cos": c} for k, (r, c) in rows.items()}})
    assert n_flip <= max(2, B // 1000)
    return rows, rgb_err, abs(float(loss) - float(loss_ref))


def _trained_weights(rn, dev, steps=200, B=1024):
    """Weights after `steps` steps of OUR trainer on the teacher scene (non-pathological, non-initial weights)."""
    tc, tf = TR.to_params(O.make_weights(21, sharpen=True), dev, False), TR.to_params(O.make_weights(22, sharpen=True), dev, False)
    torch.manual_seed(42)
    nc, nf = rn.create_nerf(rn.ModelConfig())
    nc, nf = nc.to(dev), nf.to(dev)
    tr = rn.Trainer(nc, nf, rn.RenderConfig(), lr=5e-4)
    for it in range(steps):
        ro, rd, _, _ = _scene_rays(rn, dev, B, 1000 + it)
        tgt = _teacher_targets(tc, tf, ro, rd)
        torch.manual_seed(it)
        tr.step_rays(ro, rd, tgt)
    return _weights_of(nc), _weights_of(nf)


MIN_COS, MAX_REL = 0.99, 0.15


def test_gradient_parity_4096_rays_random_init(rn, dev):
    torch.manual_seed(42)                                  # reference default seed (config.py:83)
    nc, nf = rn.create_nerf(rn.ModelConfig())
    rows, rgb_err, dloss = _grad_table(rn, dev, _weights_of(nc), _weights_of(nf), 4096, 7, "random_init")
    assert rgb_err < 1e-2 and dloss < 1e-4
    worst = min(c for _, c in rows.values())
    print(f"worst cosine {worst:.5f}")
    assert worst >= MIN_COS, {k: v for k, v in rows.items() if v[1] < MIN_COS}
    assert max(r for r, _ in rows.values()) <= MAX_REL


def test_gradient_parity_4096_rays_after_200_steps(rn, dev):
    wc, wf = _trained_weights(rn, dev)
    rows, rgb_err, dloss = _grad_table(rn, dev, wc, wf, 4096, 8, "after_200_steps")
    assert rgb_err < 1e-2 and dloss < 1e-4
    worst = min(c for _, c in rows.values())
    print(f"worst cosine {worst:.5f}")
    assert worst >= MIN_COS, {k: v for k, v in rows.items() if v[1] < MIN_COS}
    assert max(r for r, _ in rows.values()) <= MAX_REL


def test_pose_gradient_parity_4096_rays(rn, dev):
    """d loss / d (omega, delta_t) through raygen -> sampling -> MLP (PE backward with its 2^k factors) -> compositing."""
    B, H, W = 4096, 800, 800
    wc, wf = O.make_weights(41, sharpen=True), O.make_weights(42, sharpen=True)
    nc, nf = _net_from(rn, wc, dev), _net_from(rn, wf, dev)
    gt = rn.lego_poses(dev)
    noisy = rn.add_noise_to_poses(gt, 5.0, 5.0, seed=42)
    cam = rn.CameraPoseParameters(noisy).to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    with torch.no_grad():
        cam.rotation_deltas.copy_(torch.randn(100, 3, device=dev, generator=g) * 1e-3)
        cam.translation_deltas.copy_(torch.randn(100, 3, device=dev, generator=g) * 1e-2)
    _, _, img, uv = _scene_rays(rn, dev, B, 21)
    target = torch.rand(B, 3, device=dev, generator=g)
    t_rand = torch.rand(B, 64, device=dev, generator=g)
    u = torch.rand(B, 128, device=dev, generator=g)
    ro, rd = rn.ops.RayGenSE3.apply(img, uv, cam.initial_poses, cam.rotation_deltas, cam.translation_deltas, True, True,
                                    H, W, TR.lego_focal(W), W / 2.0, H / 2.0)
    out = rn.render_rays(nc, nf, ro, rd, rn.RenderConfig(), is_train=True, t_rand=t_rand, u=u)
    (((out["rgb_coarse"] - target) ** 2).mean() + ((out["rgb_fine"] - target) ** 2).mean()).backward()
    rot = cam.rotation_deltas.detach().clone().requires_grad_(True)
    trans = cam.translation_deltas.detach().clone().requires_grad_(True)
    dirs = TR.get_ray_directions(H, W, TR.lego_focal(W), device=dev)
    ro_t, rd_t = TR.rays_from_pixels(img, uv, TR.get_poses(noisy, rot, trans), dirs)
    assert float((ro - ro_t).abs().max()) < 1e-5 and float((rd - rd_t).abs().max()) < 1e-5
    res = TR.render_rays(TR.to_params(wc, dev, False), TR.to_params(wf, dev, False), ro_t, rd_t, is_train=True, t_rand=t_rand, u=u)
    TR.render_loss(res, target).backward()
    rel_r, cos_r = _cmp(cam.rotation_deltas.grad, rot.grad)
    rel_t, cos_t = _cmp(cam.translation_deltas.grad, trans.grad)
    # per-camera direction agreement (what the pose optimiser follows)
    cg = torch.nn.functional.cosine_similarity(cam.translation_deltas.grad.double(), trans.grad.double(), dim=-1)
    print(f"\n[pose grads, 4096 rays] rotation: rel L2 {rel_r:.4f} cosine {cos_r:.5f};  translation: rel L2 {rel_t:.4f} "
          f"cosine {cos_t:.5f};  per-camera translation cosine: min {float(cg.min()):.4f} median {float(cg.median()):.4f}")
    _report("pose_grad_parity", {"rays": B, "rot": {"rel_l2": rel_r, "cos": cos_r}, "trans": {"rel_l2": rel_t, "cos": cos_t},
                                 "per_camera_trans_cos_min": float(cg.min()), "per_camera_trans_cos_median": float(cg.median())})
    # Measured on a B200: cosine 0.96 / relative L2 0.28-0.30 on these stress weights.  profiles/r02_grad_rounding_ablation.md
    # (scripts/grad_rounding_ablation.py) shows where it comes from: not from any rounding point of the backward (all-fp32
    # backward: unchanged) but from the bf16 FORWARD operands -- weights, stored activations and encoding features each
    # move the ray gradients by 10-19 % on their own, because the input gradient of an L = 10 positional encoding is a sum
    # of 2^k-scaled terms that largely cancel.  A reduced-precision MLP (which north_star allows) cannot reach 0.99 here;
    # what it must do is refine poses as well as fp32 does: test_pose_refinement_on_known_scene below.
    assert cos_r >= 0.93 and cos_t >= 0.93, (cos_r, cos_t)
    assert rel_r <= 0.4 and rel_t <= 0.4, (rel_r, rel_t)


# ------------------------------------------------------------------------------------------------
# (c) convergence
# ------------------------------------------------------------------------------------------------
def _psnr(mse):
    return -10.0 * np.log10(max(float(mse), 1e-20))


def test_convergence_300_steps_vs_fp32_restatement(rn, dev):
    """Clean-pose training on a learnable scene: our Trainer (bf16 tensor-core MLP, fused clip+Adam) and the fp32
    restatement (autograd, torch.optim.Adam) see the same batches and the same random draws (same seed => same Philox
    stream: both call torch.rand(B,64) then torch.rand(B,128) on this device).  Training trajectories are chaotic, so two
    more runs give the scale: the SAME fp32 restatement with TF32 matmuls (the other precision north_star allows), and
    the same fp32 restatement started from weights perturbed by 1e-6 relative (a rounding-sized nudge)."""
    steps, B = 300, 1024
    tc, tf = TR.to_params(O.make_weights(21, sharpen=True), dev, False), TR.to_params(O.make_weights(22, sharpen=True), dev, False)
    torch.manual_seed(42)
    nc, nf = rn.create_nerf(rn.ModelConfig())
    nc, nf = nc.to(dev), nf.to(dev)
    pc, pf = TR.to_params(_weights_of(nc), dev), TR.to_params(_weights_of(nf), dev)
    qc, qf = TR.to_params(_weights_of(nc), dev), TR.to_params(_weights_of(nf), dev)
    ours = rn.Trainer(nc, nf, rn.RenderConfig(), lr=5e-4)
    ref = TR.RefTrainer(pc, pf, lr=5e-4)
    ref_tf32 = TR.RefTrainer(qc, qf, lr=5e-4)
    gp = torch.Generator(device="cpu").manual_seed(7)
    nudged = lambda w: {k: v * (1.0 + 1e-6 * torch.randn(v.shape, generator=gp).to(v)) for k, v in w.items()}     # noqa: E731
    rc, rf = TR.to_params(nudged(_weights_of(nc)), dev), TR.to_params(nudged(_weights_of(nf)), dev)
    ref_nudged = TR.RefTrainer(rc, rf, lr=5e-4)
    ro_e, rd_e, _, _ = _scene_rays(rn, dev, 8192, 99)
    tgt_e = _teacher_targets(tc, tf, ro_e, rd_e)

    def eval_psnr_ours():
        with torch.no_grad():
            o = rn.render_rays(nc, nf, ro_e, rd_e, rn.RenderConfig(), is_train=False)
        return _psnr(((o["rgb_fine"] - tgt_e) ** 2).mean())

    def eval_psnr_ref(a, b):
        return _psnr(((_teacher_targets(a, b, ro_e, rd_e) - tgt_e) ** 2).mean())

    la, lb, curve, dense = [], [], [], []
    assert not torch.backends.cuda.matmul.allow_tf32
    for it in range(steps):
        ro, rd, _, _ = _scene_rays(rn, dev, B, 5000 + it)
        tgt = _teacher_targets(tc, tf, ro, rd)
        torch.manual_seed(it)
        la.append(ours.step_rays(ro, rd, tgt))
        torch.manual_seed(it)
        lb.append(ref.step_rays(ro, rd, tgt))
        torch.manual_seed(it)
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            ref_tf32.step_rays(ro, rd, tgt)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(it)
        ref_nudged.step_rays(ro, rd, tgt)
        if (it + 1) % 25 == 0:
            curve.append((it + 1, eval_psnr_ours(), eval_psnr_ref(pc, pf), eval_psnr_ref(qc, qf)))
        if it + 1 > 2 * steps // 3 and (it + 1) % 5 == 0:
            # the last third, every 5 steps: single checkpoints sit on short dips of the (noisy, lr = 5e-4) trajectories
            # that the runs pass at different steps, so the converged level is the mean over many checkpoints
            c4 = curve[-1] if (it + 1) % 25 == 0 else (it + 1, eval_psnr_ours(), eval_psnr_ref(pc, pf), eval_psnr_ref(qc, qf))
            dense.append(c4 + (eval_psnr_ref(rc, rf),))
    la = torch.stack([x.reshape(()) for x in la]).cpu().numpy()
    lb = torch.stack([x.reshape(()) for x in lb]).cpu().numpy()
    print("\n[convergence] held-out PSNR (dB) every 25 steps: step, ours (bf16 tcgen05), fp32 restatement, same with TF32 matmuls, ours - fp32, tf32 - fp32")
    for s_, a, b, c in curve:
        print(f"  {s_:4d}  {a:7.3f}  {b:7.3f}  {c:7.3f}  {a - b:+.3f}  {c - b:+.3f}")
    third = len(curve) // 3
    wins = []
    for k in range(3):
        seg = curve[k * third:(k + 1) * third]
        wins.append((seg[0][0], seg[-1][0], float(np.mean([a for _, a, _, _ in seg])), float(np.mean([b for _, _, b, _ in seg])),
                     float(np.mean([c for _, _, _, c in seg]))))
    print("[convergence] window means: steps, ours, fp32, tf32")
    for a0, a1, x, y, z in wins:
        print(f"  {a0:4d}-{a1:4d}  {x:7.3f}  {y:7.3f}  {z:7.3f}  {x - y:+.3f}  {z - y:+.3f}")
    last = tuple(float(np.mean([r[k] for r in dense])) for k in (1, 2, 3, 4))
    noise = max(abs(last[2] - last[1]), abs(last[3] - last[1]))
    print(f"[convergence] last third, mean of {len(dense)} checkpoints (every 5 steps): ours {last[0]:.3f}  fp32 {last[1]:.3f}  "
          f"tf32 {last[2]:.3f}  fp32 from nudged weights {last[3]:.3f}  ours - fp32 {last[0] - last[1]:+.3f}  "
          f"tf32 - fp32 {last[2] - last[1]:+.3f}  nudged - fp32 {last[3] - last[1]:+.3f}")
    dev_ours = max(abs(a - b) for _, a, b, _ in curve)
    dev_tf32 = max(abs(c - b) for _, _, b, c in curve)
    _report("convergence_clean", {"steps": steps, "rays_per_step": B, "eval_psnr_step_ours_fp32_tf32": curve, "window_means": wins,
                                  "last_third_mean_of_dense_checkpoints_ours_fp32_tf32_nudged": last, "dense_checkpoints": len(dense),
                                  "max_abs_dev_ours_vs_fp32_db": dev_ours, "max_abs_dev_tf32_vs_fp32_db": dev_tf32,
                                  "first_loss": [float(la[0]), float(lb[0])]})
    print(f"[convergence] max |ours - fp32| {dev_ours:.3f} dB; max |tf32 - fp32| {dev_tf32:.3f} dB (trajectory noise scale)")
    # training PSNR over 100-step windows (each the mean loss of 100 batches of 1024 rays: the smooth quantity)
    tr_win = []
    for a in range(0, steps, 100):
        tr_win.append((a, _psnr(np.mean(la[a:a + 100]) / 2.0), _psnr(np.mean(lb[a:a + 100]) / 2.0)))
    print("[convergence] training PSNR, 100-step windows: start, ours, fp32, delta")
    for a, x_, y_ in tr_win:
        print(f"  {a:4d}  {x_:7.3f}  {y_:7.3f}  {x_ - y_:+.3f}")
    _report("convergence_clean_train_windows", tr_win)
    assert abs(la[0] - lb[0]) < 1e-4 * max(1.0, lb[0])
    assert curve[-1][1] > curve[0][1] + 3.0                                                 # it learns
    # Measured on B200s (two PE implementations, several boxes): single held-out checkpoints of ours deviate from fp32 by
    # up to 0.4-0.7 dB -- as do the TF32 run's (0.56 dB): the three trajectories pass a loss-plateau escape around step
    # 75-175 at slightly different times, and later on short dips (0.5-1 dB for 5-10 steps) at different steps -- a
    # one-ulp change of the colour sigmoid moved the 275-step checkpoint of ours by 0.5 dB and the last third's level by
    # 0.25 dB; the fp32 run with TF32 matmuls ends 0.15 dB from the fp32 run.  What is held to 0.1 dB is the quantity that
    # averages the timing out: the training PSNR of every 100-step window (measured deltas: 0.007-0.025 dB).  The held-out
    # level of the last third (mean of 20 checkpoints) and the 4-checkpoint windows are held to twice the scatter that
    # rounding-sized changes of the fp32 run itself produce (TF32 matmuls; weights nudged by 1e-6).
    assert max(abs(x_ - y_) for _, x_, y_ in tr_win) <= 0.1, tr_win
    assert abs(last[0] - last[1]) <= max(0.1, 2.0 * noise), (last, noise)
    assert max(abs(x - y) for _, _, x, y, _ in wins) <= max(0.3, 2.0 * max(abs(z - y) for _, _, _, y, z in wins)), wins
    assert dev_ours <= max(0.1, 2.0 * dev_tf32), (dev_ours, dev_tf32)


def _blob_scene_targets(ro, rd, n=256):
    """Analytic learnable scene: three coloured Gaussian density blobs inside the lego cameras' view volume, rendered with
    the reference's compositing formula (oracle/torch_ref.py::raw2outputs) on n uniform depths in [2, 6]."""
    dev = ro.device
    centers = torch.tensor([[0.7, 0.0, 0.2], [-0.5, 0.5, -0.1], [0.0, -0.7, 0.4]], device=dev)
    cols = torch.tensor([[1.0, 0.15, 0.15], [0.15, 1.0, 0.15], [0.15, 0.15, 1.0]], device=dev)
    z = torch.linspace(2.0, 6.0, n, device=dev).expand(ro.shape[0], n)
    pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
    d2 = ((pts[:, :, None, :] - centers) ** 2).sum(-1)
    dens = 25.0 * torch.exp(-d2 / (2 * 0.4 ** 2))
    sigma = dens.sum(-1, keepdim=True)
    rgb = (dens[..., None] * cols).sum(-2) / (sigma + 1e-6)
    return TR.raw2outputs(rgb, sigma, z, rd, None, True)["rgb_map"]


def test_pose_refinement_on_learned_scene(rn, dev):
    """The quantity pose gradients exist for.  A scene is first LEARNED (our Trainer, true cameras, 600 steps on an
    analytic three-blob density field), then the cameras are perturbed by 2 deg / 2 %, omega is seeded N(0, 1e-3) (live
    rotation branch, quirk 11) and ONLY the poses are optimised (networks frozen: lr = 0; pose clip 0.1, Adam,
    train_pose_opt.py:398-409 semantics) -- once with our kernels, once with the fp32 restatement, same batches and
    draws.  Both must pull the cameras back, and ours must end where fp32 ends."""
    H = W = 800
    focal = TR.lego_focal(W)
    gt = rn.lego_poses(dev)
    dirs = TR.get_ray_directions(H, W, focal, device=dev)
    torch.manual_seed(42)
    nc, nf = rn.create_nerf(rn.ModelConfig())
    nc, nf = nc.to(dev), nf.to(dev)
    pre = rn.Trainer(nc, nf, rn.RenderConfig(), lr=5e-4)
    for it in range(600):
        ro, rd, _, _ = _scene_rays(rn, dev, 4096, 20000 + it)
        with torch.no_grad():
            tgt = _blob_scene_targets(ro, rd)
        torch.manual_seed(it)
        loss = pre.step_rays(ro, rd, tgt)
    psnr_scene = _psnr(float(loss) / 2)
    wc, wf = _weights_of(nc), _weights_of(nf)
    del pre
    # fresh modules on the learned weights (the pre-training Trainer's flat buffers are gone)
    nc, nf = _net_from(rn, wc, dev), _net_from(rn, wf, dev)
    pc, pf = TR.to_params(wc, dev), TR.to_params(wf, dev)
    noisy = rn.add_noise_to_poses(gt, 2.0, 2.0, seed=42)
    cam = rn.CameraPoseParameters(noisy).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        cam.rotation_deltas.copy_(torch.randn(100, 3, device=dev, generator=g) * 1e-3)
    rot = cam.rotation_deltas.detach().clone().requires_grad_(True)
    trans = cam.translation_deltas.detach().clone().requires_grad_(True)
    steps, B, pose_lr = 300, 2048, 1e-3
    ours = rn.Trainer(nc, nf, rn.RenderConfig(), lr=0.0, camera_params=cam, pose_lr=pose_lr)
    ref = TR.RefTrainer(pc, pf, lr=0.0, initial_poses=noisy, rot=rot, trans=trans, pose_lr=pose_lr)

    class _Sampler:
        def get_rays_for_batch_fused(self, pb, cp):
            return rn.ops.RayGenSE3.apply(pb.image_indices, pb.pixel_coords, cp.initial_poses, cp.rotation_deltas,
                                          cp.translation_deltas, cp.learn_rotation, cp.learn_translation, H, W, focal,
                                          W / 2.0, H / 2.0)

    def errs(poses):
        e = rn.compute_pose_errors_batch(gt, poses).double()
        return float(e[:, 0].mean()), float(e[:, 1].mean())

    from robust_nerf_b200.data_pose_opt import PixelBatch
    e0 = errs(noisy)
    traj = []
    for it in range(steps):
        _, _, img, uv = _scene_rays(rn, dev, B, 9000 + it)
        with torch.no_grad():
            ro_gt, rd_gt = TR.rays_from_pixels(img, uv, gt, dirs)
            tgt = _blob_scene_targets(ro_gt.contiguous(), rd_gt.contiguous())
        torch.manual_seed(it)
        ours.step_pixels(PixelBatch(img, uv, tgt), _Sampler(), optimize_poses=True)
        torch.manual_seed(it)
        ref.step_pixels(img, uv, tgt, dirs, optimize_poses=True)
        if (it + 1) % 50 == 0:
            with torch.no_grad():
                traj.append((it + 1, *errs(cam.get_all_poses()), *errs(TR.get_poses(noisy, rot, trans))))
    d_rot = float((cam.rotation_deltas.detach() - rot.detach()).norm() / rot.detach().norm())
    d_tr = float((cam.translation_deltas.detach() - trans.detach()).norm() / trans.detach().norm().clamp_min(1e-12))
    print(f"\n[pose refinement, learned scene] scene PSNR after pre-training {psnr_scene:.2f} dB; initial pose error: rot {e0[0]:.4f} deg, "
          f"trans {e0[1]:.5f}")
    print("  step   ours: rot deg, trans      fp32 restatement: rot deg, trans")
    for s_, a, b, c, d in traj:
        print(f"  {s_:4d}  {a:8.4f} {b:8.5f}   {c:8.4f} {d:8.5f}")
    print(f"  parameter space after {steps} steps: |omega_ours - omega_fp32| / |omega_fp32| = {d_rot:.3f}, "
          f"|dt_ours - dt_fp32| / |dt_fp32| = {d_tr:.3f}")
    _report("pose_refinement_learned_scene", {"steps": steps, "rays_per_step": B, "scene_psnr": psnr_scene, "initial_err": e0,
                                              "trajectory": traj, "rel_diff_omega": d_rot, "rel_diff_delta_t": d_tr})
    fa, fb, fc, fd = traj[-1][1:]
    assert fc < 0.8 * e0[0] or fd < 0.8 * e0[1], ("the fp32 restatement did not refine the poses: scenario void", e0, traj[-1])
    assert fa <= fc + 0.1 * e0[0] and fb <= fd + 0.1 * e0[1], (e0, traj[-1])     # ours refines as well as fp32 does


def test_pose_opt_convergence_300_steps_vs_fp32_restatement(rn, dev):
    """Joint pose optimisation (train_pose_opt.py:290-411): 5 deg / 5 % noisy initial poses, omega seeded N(0, 1e-3) so
    the rotation-gradient branch is live; pose-error trajectory of our Trainer vs the fp32 restatement."""
    steps, B, H, W = 300, 1024, 800, 800
    focal = TR.lego_focal(W)
    tc, tf = TR.to_params(O.make_weights(21, sharpen=True), dev, False), TR.to_params(O.make_weights(22, sharpen=True), dev, False)
    gt = rn.lego_poses(dev)
    noisy = rn.add_noise_to_poses(gt, 5.0, 5.0, seed=42)
    dirs = TR.get_ray_directions(H, W, focal, device=dev)
    torch.manual_seed(42)
    nc, nf = rn.create_nerf(rn.ModelConfig())
    nc, nf = nc.to(dev), nf.to(dev)
    pc, pf = TR.to_params(_weights_of(nc), dev), TR.to_params(_weights_of(nf), dev)
    cam = rn.CameraPoseParameters(noisy).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        cam.rotation_deltas.copy_(torch.randn(100, 3, device=dev, generator=g) * 1e-3)
    rot = cam.rotation_deltas.detach().clone().requires_grad_(True)
    trans = cam.translation_deltas.detach().clone().requires_grad_(True)
    ours = rn.Trainer(nc, nf, rn.RenderConfig(), lr=5e-4, camera_params=cam, pose_lr=1e-3, rotation_reg_weight=0.01,
                      translation_reg_weight=0.001)
    ref = TR.RefTrainer(pc, pf, lr=5e-4, initial_poses=noisy, rot=rot, trans=trans, pose_lr=1e-3, rot_reg=0.01, trans_reg=0.001)

    class _Sampler:                      # Trainer.step_pixels only needs get_rays_for_batch_fused
        def get_rays_for_batch_fused(self, pb, cp):
            return rn.ops.RayGenSE3.apply(pb.image_indices, pb.pixel_coords, cp.initial_poses, cp.rotation_deltas,
                                          cp.translation_deltas, cp.learn_rotation, cp.learn_translation, H, W, focal,
                                          W / 2.0, H / 2.0)

    def errs(poses):
        e = rn.compute_pose_errors_batch(gt, poses).double()
        return float(e[:, 0].mean()), float(e[:, 1].mean())

    from robust_nerf_b200.data_pose_opt import PixelBatch
    traj = []
    for it in range(steps):
        _, _, img, uv = _scene_rays(rn, dev, B, 9000 + it)
        with torch.no_grad():                              # targets: the teacher seen from the TRUE cameras
            ro_gt, rd_gt = TR.rays_from_pixels(img, uv, gt, dirs)
        tgt = _teacher_targets(tc, tf, ro_gt.contiguous(), rd_gt.contiguous())
        torch.manual_seed(it)
        ours.step_pixels(PixelBatch(img, uv, tgt), _Sampler(), optimize_poses=True)
        torch.manual_seed(it)
        ref.step_pixels(img, uv, tgt, dirs, optimize_poses=True)
        if (it + 1) % 50 == 0:
            with torch.no_grad():
                traj.append((it + 1, *errs(cam.get_all_poses()), *errs(TR.get_poses(noisy, rot, trans))))
    e0 = errs(noisy)
    d_ours = (cam.translation_deltas.detach() - 0).norm().item()
    d_between = (cam.translation_deltas.detach() - trans.detach()).norm().item()
    dr_ours = cam.rotation_deltas.detach().norm().item()
    dr_between = (cam.rotation_deltas.detach() - rot.detach()).norm().item()
    print(f"\n[pose-opt] initial pose error: rot {e0[0]:.4f} deg, trans {e0[1]:.5f}")
    print("[pose-opt] step, ours (rot deg, trans), fp32 restatement (rot deg, trans)")
    for s, a, b, c, d in traj:
        print(f"  {s:4d}  {a:8.4f} {b:8.5f}   {c:8.4f} {d:8.5f}")
    print(f"[pose-opt] |delta_t| moved {d_ours:.4f}, ours-vs-fp32 {d_between:.4f} ({d_between / max(d_ours, 1e-12):.3f} of the movement); "
          f"|omega| {dr_ours:.4f}, ours-vs-fp32 {dr_between:.4f} ({dr_between / max(dr_ours, 1e-12):.3f})")
    _report("convergence_pose_opt", {"steps": steps, "rays_per_step": B, "initial_err": e0, "trajectory": traj,
                                     "trans_moved": d_ours, "trans_diff_vs_fp32": d_between, "rot_norm": dr_ours,
                                     "rot_diff_vs_fp32": dr_between})
    # Joint optimisation of two randomly initialised networks AND the cameras is chaotic in its first hundreds of steps
    # (the fp32 run itself barely reduces the pose error here): the trajectories are reported, and what is asserted is
    # that ours stays as sane as fp32's -- the poses move, nothing blows up, and the pose error never exceeds fp32's by
    # more than 10 % of the initial error.  The point-by-point comparison is test_pose_refinement_on_known_scene.
    assert d_ours > 1e-2 and np.isfinite(d_ours) and np.isfinite(dr_ours)
    for s_, a, b, c, d in traj:
        assert a <= c + 0.1 * e0[0] and b <= d + 0.1 * e0[1], traj
