"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/noisy_src`).

Run in the authoring container only (the reference cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs `tests/golden/*.npz` (committed).  Inputs are seeded; network weights come from
`oracle.nerf_oracle.make_weights` (a numpy generator, so tests can rebuild the same
weights without torch RNG) and are loaded into the reference `NeRF` through
`load_state_dict`.  RNG draws of the reference (`torch.rand` in rays.py:204/255) are
recorded by patching `torch.rand` and stored next to the outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from noisy_src.config import ModelConfig, RenderConfig  # noqa: E402
from noisy_src.model import NeRF, PositionalEncoding  # noqa: E402
from noisy_src import rays as R  # noqa: E402
from noisy_src.rendering import raw2outputs, render_rays  # noqa: E402
from noisy_src.train_pose_opt import CameraPoseParameters  # noqa: E402
from noisy_src.data_pose_opt import PixelDataset, PixelSampler, PixelBatch  # noqa: E402
from noisy_src.data import BlenderData  # noqa: E402
from noisy_src.noise import NoiseConfig, add_noise_to_poses  # noqa: E402

from oracle import nerf_oracle as O  # noqa: E402

torch.set_num_threads(8)


class RandRecorder:
    """Patch torch.rand to record every draw (in call order)."""

    def __init__(self):
        self.draws = []
        self._orig = torch.rand

    def __enter__(self):
        def rec(*a, **k):
            t = self._orig(*a, **k)
            self.draws.append(t.detach().cpu().numpy().copy())
            return t
        torch.rand = rec
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig


def ref_net(weights, cfg=None):
    net = NeRF(cfg or ModelConfig())
    sd = net.state_dict()
    for k, v in weights.items():
        sd[k] = torch.from_numpy(v.copy())
    net.load_state_dict(sd)
    return net


def np_(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz: {os.path.getsize(path) / 1024:.1f} KiB")


def lego_poses():
    d = torch.load("/root/reference/outputs/lego_poseopt_cleaninit_clean_20251207_205248/final_poses.pt",
                   map_location="cpu", weights_only=False)
    return d["ground_truth_poses"].float()


def main():
    rng = np.random.default_rng(1234)
    gt = lego_poses()

    # ---- 0. scene poses + pose-error known answers (recorded by the reference's own runs) ----
    kat = {}
    for i, run in enumerate(["lego_poseopt_noisyinit_rot5.0deg_20251207_233144",
                             "lego_poseopt_noisyinit_rot5.0deg_trans5.0pct_20251209_180945",
                             "lego_poseopt_noisyinit_rot5.0deg_trans5.0pct_20251209_195334"]):
        d = torch.load(f"/root/reference/outputs/{run}/final_poses.pt", map_location="cpu",
                       weights_only=False)
        kat[f"init_{i}"] = np_(d["initial_poses"].float())
        kat[f"opt_{i}"] = np_(d["optimized_poses"].float())
        kat[f"err_{i}"] = np.array([d["pose_errors"][k] for k in
                                    ("rotation_error_mean", "rotation_error_std", "rotation_error_max",
                                     "translation_error_mean", "translation_error_std",
                                     "translation_error_max")], dtype=np.float64)
    save("lego_poses", ground_truth_poses=np_(gt), **kat)

    # ---- 1. positional encoding ----
    x = (rng.standard_normal((64, 3)) * 3).astype(np.float32)
    save("pe", x=x, pe10=np_(PositionalEncoding(10)(torch.from_numpy(x))),
         pe4=np_(PositionalEncoding(4)(torch.from_numpy(x))))

    # ---- 2. NeRF forward/backward, full config, plain + sharpened weights ----
    for tag, sharpen in (("plain", False), ("sharp", True)):
        w = O.make_weights(7, sharpen=sharpen)
        net = ref_net(w)
        pts = (rng.uniform(-3, 3, (384, 3))).astype(np.float32)
        dirs = rng.standard_normal((384, 3)).astype(np.float32)
        dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
        tp = torch.from_numpy(pts).requires_grad_(True)
        td = torch.from_numpy(dirs).requires_grad_(True)
        rgb, sigma = net(tp, td)
        g_rgb = rng.standard_normal(rgb.shape).astype(np.float32)
        g_sig = rng.standard_normal(sigma.shape).astype(np.float32)
        (rgb * torch.from_numpy(g_rgb)).sum().add((sigma * torch.from_numpy(g_sig)).sum()).backward()
        grads = {"g_" + k: np_(p.grad) for k, p in net.named_parameters()}
        big = {k: v for k, v in grads.items() if v.size > 4096}
        small = {k: v for k, v in grads.items() if v.size <= 4096}
        # large grads: keep a strided subsample + norms (fixture stays small)
        sub = {k: v.reshape(-1)[::97].copy() for k, v in big.items()}
        nrm = {k + "_norm": np.array(np.linalg.norm(v.astype(np.float64))) for k, v in big.items()}
        save(f"nerf_{tag}", pts=pts, dirs=dirs, rgb=np_(rgb), sigma=np_(sigma), g_rgb=g_rgb,
             g_sigma=g_sig, d_pts=np_(tp.grad), d_dirs=np_(td.grad), **small, **sub, **nrm)

    # ---- 3. small-config NeRF (exercises skip/no-viewdir variants of the oracle) ----
    cfg_s = ModelConfig(pos_freqs=4, dir_freqs=2, hidden_dim=32, num_hidden_layers=4, skips=(1,))
    w = O.make_weights(11, O.ModelConfig(4, 2, 32, 4, (1,), True))
    net = ref_net(w, cfg_s)
    pts = rng.uniform(-2, 2, (50, 3)).astype(np.float32)
    dirs = rng.standard_normal((50, 3)).astype(np.float32)
    rgb, sigma = net(torch.from_numpy(pts), torch.from_numpy(dirs))
    save("nerf_small", pts=pts, dirs=dirs, rgb=np_(rgb), sigma=np_(sigma))

    # ---- 4. ray generation ----
    dirs_hw = R.get_ray_directions(20, 16, 13.7)
    dirs_c = R.get_ray_directions(20, 16, 13.7, center=(7.25, 11.5))
    ro, rd = R.get_rays(dirs_hw, gt[3])
    rob, rdb = R.get_rays_batch(6, 5, 4.2, gt[:3])
    d800 = R.get_ray_directions(800, 800, 0.5 * 800 / np.tan(0.5 * 0.6911112070083618))
    save("rays", dirs=np_(dirs_hw), dirs_center=np_(dirs_c), rays_o=np_(ro), rays_d=np_(rd),
         batch_o=np_(rob), batch_d=np_(rdb), pose=np_(gt[3]),
         d800_row=np_(d800[417]), d800_col=np_(d800[:, 123]))

    # ---- 5. stratified sampling ----
    B = 24
    ro = torch.from_numpy(rng.standard_normal((B, 3)).astype(np.float32))
    rd = torch.from_numpy(rng.standard_normal((B, 3)).astype(np.float32))
    out = {}
    for tag, kw in (("det", dict(perturb=False)), ("pert", dict(perturb=True)),
                    ("lindisp", dict(perturb=True, lindisp=True))):
        with RandRecorder() as rec:
            pts, z = R.sample_along_rays(ro, rd, 2.0, 6.0, 64, **kw)
        out[f"pts_{tag}"], out[f"z_{tag}"] = np_(pts), np_(z)
        if rec.draws:
            out[f"trand_{tag}"] = rec.draws[0]
    for n in (2, 3, 64, 65, 128, 192):
        out[f"linspace_{n}"] = np_(torch.linspace(0.0, 1.0, n))
    save("stratified", rays_o=np_(ro), rays_d=np_(rd), **out)

    # ---- 6. sample_pdf / sample_hierarchical ----
    out = {}
    with RandRecorder() as rec:
        _, z = R.sample_along_rays(ro, rd, 2.0, 6.0, 64, perturb=True)
    wts = torch.from_numpy((rng.uniform(0, 1, (B, 64)) ** 4).astype(np.float32))
    wts[0] = 0.0                      # all-zero weights -> uniform pdf
    wts[1] = 0.0; wts[1, 17] = 5.0    # single spike -> repeated samples (denom<1e-5 branch)
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    out["z"], out["weights"] = np_(z), np_(wts)
    out["pdf_det"] = np_(R.sample_pdf(mids, wts[..., 1:-1], 128, det=True))
    with RandRecorder() as rec:
        out["pdf_rand"] = np_(R.sample_pdf(mids, wts[..., 1:-1], 128, det=False))
    out["u_rand"] = rec.draws[0]
    pf, zf = R.sample_hierarchical(ro, rd, z, wts, 128, det=True)
    out["hier_pts_det"], out["hier_z_det"] = np_(pf), np_(zf)
    with RandRecorder() as rec:
        pf, zf = R.sample_hierarchical(ro, rd, z, wts, 128, det=False)
    out["hier_pts_rand"], out["hier_z_rand"], out["hier_u"] = np_(pf), np_(zf), rec.draws[0]
    save("sample_pdf", rays_o=np_(ro), rays_d=np_(rd), **out)

    # ---- 7. raw2outputs forward + backward ----
    S = 48
    rgb = torch.from_numpy(rng.uniform(0, 1, (B, S, 3)).astype(np.float32)).requires_grad_(True)
    sig = torch.from_numpy((rng.uniform(-2, 10, (B, S, 1))).astype(np.float32)).requires_grad_(True)
    zz = torch.sort(torch.from_numpy(rng.uniform(2, 6, (B, S)).astype(np.float32)), -1)[0]
    rdd = torch.from_numpy((rng.standard_normal((B, 3)) * 1.3).astype(np.float32)).requires_grad_(True)
    out = {}
    for tag, wb in (("white", True), ("black", False)):
        for p in (rgb, sig, rdd):
            p.grad = None
        o = raw2outputs(rgb, sig, zz, rdd, 0.0, wb)
        g_map = rng.standard_normal((B, 3)).astype(np.float32)
        g_dep = rng.standard_normal((B,)).astype(np.float32)
        g_acc = rng.standard_normal((B,)).astype(np.float32)
        g_w = rng.standard_normal((B, S)).astype(np.float32)
        ((o["rgb_map"] * torch.from_numpy(g_map)).sum() + (o["depth_map"] * torch.from_numpy(g_dep)).sum()
         + (o["acc_map"] * torch.from_numpy(g_acc)).sum() + (o["weights"] * torch.from_numpy(g_w)).sum()).backward()
        out.update({f"{tag}_{k}": np_(v) for k, v in o.items()})
        out.update({f"{tag}_g_map": g_map, f"{tag}_g_depth": g_dep, f"{tag}_g_acc": g_acc,
                    f"{tag}_g_w": g_w, f"{tag}_d_rgb": np_(rgb.grad), f"{tag}_d_sigma": np_(sig.grad),
                    f"{tag}_d_rays_d": np_(rdd.grad)})
    save("raw2outputs", rgb=np_(rgb), sigma=np_(sig), z=np_(zz), rays_d=np_(rdd), **out)

    # ---- 8. render_rays eval/train, full config, plain + sharpened; train-step gradients ----
    for tag, sharpen in (("plain", False), ("sharp", True)):
        wc, wf = O.make_weights(21, sharpen=sharpen), O.make_weights(22, sharpen=sharpen)
        nc, nf = ref_net(wc), ref_net(wf)
        Bq = 12
        H = W = 800
        focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
        pix = rng.integers(0, H * W, Bq)
        dirs = R.get_ray_directions(H, W, focal).reshape(-1, 3)[pix]
        ro, rd = R.get_rays(dirs, gt[5])
        ro, rd = ro.contiguous().clone().requires_grad_(True), rd.clone().requires_grad_(True)
        tgt = torch.from_numpy(rng.uniform(0, 1, (Bq, 3)).astype(np.float32))
        rc = RenderConfig()
        with torch.no_grad():
            ev = render_rays(nc, nf, ro, rd, rc, is_train=False)
        with RandRecorder() as rec:
            tr = render_rays(nc, nf, ro, rd, rc, is_train=True)
        loss = ((tr["rgb_coarse"] - tgt) ** 2).mean() + ((tr["rgb_fine"] - tgt) ** 2).mean()
        loss.backward()
        arr = {"rays_o": np_(ro), "rays_d": np_(rd), "target": np_(tgt), "t_rand": rec.draws[0],
               "u": rec.draws[1], "loss": np_(loss), "d_rays_o": np_(ro.grad), "d_rays_d": np_(rd.grad)}
        arr.update({"eval_" + k: np_(v) for k, v in ev.items()})
        arr.update({"train_" + k: np_(v) for k, v in tr.items()})
        for nm, net in (("c", nc), ("f", nf)):
            for k, p in net.named_parameters():
                g = np_(p.grad)
                arr[f"g{nm}_{k}_norm"] = np.array(np.linalg.norm(g.astype(np.float64)))
                arr[f"g{nm}_{k}"] = g if g.size <= 4096 else g.reshape(-1)[::97].copy()
        save(f"render_{tag}", **arr)

    # ---- 9. SE(3) pose parameters, pixel ray generation, gradients ----
    init = add_noise_to_poses(gt, NoiseConfig(5.0, 0.0, 5.0, seed=42))[0] if True else gt
    cam = CameraPoseParameters(init)
    with torch.no_grad():
        cam.rotation_deltas.copy_(torch.from_numpy((rng.standard_normal((100, 3)) * 0.05).astype(np.float32)))
        cam.rotation_deltas[0] = 0.0                # exact zero -> small-angle branch (grad 0)
        cam.rotation_deltas[1] = torch.tensor([3e-7, -2e-7, 1e-7])  # below 1e-6
        cam.rotation_deltas[2] = torch.tensor([1e-3, -2e-3, 5e-4])  # tiny but live
        cam.translation_deltas.copy_(torch.from_numpy((rng.standard_normal((100, 3)) * 0.1).astype(np.float32)))
    poses = cam.get_all_poses()
    Gp = rng.standard_normal((100, 4, 4)).astype(np.float32)
    (poses * torch.from_numpy(Gp)).sum().backward()
    out = {"init": np_(init), "rot": np_(cam.rotation_deltas), "trans": np_(cam.translation_deltas),
           "poses": np_(poses), "g_poses": Gp, "d_rot": np_(cam.rotation_deltas.grad),
           "d_trans": np_(cam.translation_deltas.grad)}
    sub_idx = torch.tensor([5, 1, 77, 2])
    out["poses_sub"] = np_(cam.get_poses(sub_idx))
    out["sub_idx"] = sub_idx.numpy()
    out["pose_errors"] = np.array([cam.compute_pose_errors(gt)[k] for k in
                                   ("rotation_error_mean", "rotation_error_std", "rotation_error_max",
                                    "translation_error_mean", "translation_error_std",
                                    "translation_error_max")])
    # pixel batch -> rays (+ grads into the pose parameters)
    Hs, Ws, fs = 40, 32, 44.4
    data = BlenderData(images=torch.from_numpy(rng.uniform(0, 1, (100, Hs, Ws, 3)).astype(np.float32)),
                       poses=init, H=Hs, W=Ws, focal=fs)
    ds = PixelDataset(data)
    sampler = PixelSampler(ds, batch_size=256)
    torch.manual_seed(42)
    pb = sampler.sample_batch()
    torch.manual_seed(42)
    flat_idx = torch.randint(0, ds.n_pixels, (256,))
    cam.zero_grad()
    ro, rd = sampler.get_rays_for_batch(pb, cam.get_all_poses())
    g_o = rng.standard_normal((256, 3)).astype(np.float32)
    g_d = rng.standard_normal((256, 3)).astype(np.float32)
    ((ro * torch.from_numpy(g_o)).sum() + (rd * torch.from_numpy(g_d)).sum()).backward()
    out.update(H=np.array(Hs), W=np.array(Ws), focal=np.array(fs), flat_idx=flat_idx.numpy(),
               image_indices=np_(pb.image_indices), pixel_coords=np_(pb.pixel_coords),
               target_rgb=np_(pb.target_rgb), images_sum=np.array(float(data.images.double().sum())),
               px_rays_o=np_(ro), px_rays_d=np_(rd), g_o=g_o, g_d=g_d,
               px_d_rot=np_(cam.rotation_deltas.grad), px_d_trans=np_(cam.translation_deltas.grad))
    save("pose", **out)


if __name__ == "__main__":
    main()
