#!/usr/bin/env python
"""Where does the distance between the bf16 tensor-core path and the fp32 reference gradients come from?

CPU experiment on the numpy oracle (oracle/nerf_oracle.py), whose `emulate_bf16` mode rounds exactly where the CUDA
kernels round (the kernels agree with it to <= 1.2 %, tests/test_gpu_parity.py).  Each row switches groups of rounding
points off (oracle.BF16_POINTS) and reports relative L2 and cosine of d loss / d rays_o, d loss / d rays_d (what the pose
gradient is assembled from) and of the fine network's first-layer weight gradient against the all-fp32 oracle.

    python scripts/grad_rounding_ablation.py [rays] > profiles/r02_grad_rounding_ablation.md
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
poses = np.load(os.path.join(ROOT, "robust-nerf_b200", "data", "lego_train_poses.npy"))
H = W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
img = rng.integers(0, 100, B)
uv = np.stack([rng.integers(0, W, B), rng.integers(0, H, B)], -1).astype(np.float32)
ro, rd = O.get_rays_from_pixels(img, uv, poses, H, W, focal)
target = rng.uniform(0, 1, (B, 3)).astype(np.float32)
t_rand = rng.uniform(0, 1, (B, 64)).astype(np.float32)
u = rng.uniform(0, 1, (B, 128)).astype(np.float32)


def cmp(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)), float(a @ b / np.linalg.norm(a) / np.linalg.norm(b))


ALL = dict(w=True, act=True, enc=True, dh=True, dxe=True, bw=True)
ROWS = [
    ("all rounding points (the kernels as built)", ALL),
    ("encoding-gradient outputs dXE0/dXE5/dDE kept fp32", dict(ALL, dxe=False)),
    ("+ back-propagated activation gradients fp32", dict(ALL, dxe=False, dh=False)),
    ("whole backward fp32 (forward bf16 only)", dict(ALL, dxe=False, dh=False, bw=False)),
    ("forward: weights only (backward fp32)", dict(w=True, act=False, enc=False, dh=False, dxe=False, bw=False)),
    ("forward: stored activations only (backward fp32)", dict(w=False, act=True, enc=False, dh=False, dxe=False, bw=False)),
    ("forward: encoding features only (backward fp32)", dict(w=False, act=False, enc=True, dh=False, dxe=False, bw=False)),
]
print(f"# bf16 rounding-point ablation ({B} rays, 64+128 samples, numpy oracle, CPU)\n")
print("rel = relative L2 distance to the all-fp32 oracle, cos = cosine with it.\n")
for label, sharpen in (("random-init-like weights (nn.Linear default init)", False), ("sharpened weights (sigma x300, rgb x30: the parity tests' stress weights)", True)):
    wc, wf = O.make_weights(41, sharpen=sharpen), O.make_weights(42, sharpen=sharpen)
    O.BF16_POINTS.update(ALL)
    ref = O.train_step_grads(wc, wf, ro, rd, target, t_rand=t_rand, u=u, need_ray_grad=True)
    print(f"## {label}\n")
    print("| rounding points active | d rays_o rel / cos | d rays_d rel / cos | fine pts_linears.0.weight rel / cos | rgb_fine max-abs |")
    print("|---|---|---|---|---|")
    for name, flags in ROWS:
        O.BF16_POINTS.update(flags)
        e = O.train_step_grads(wc, wf, ro, rd, target, t_rand=t_rand, u=u, need_ray_grad=True, emulate_bf16=True)
        a, b, c = cmp(e["d_rays_o"], ref["d_rays_o"]), cmp(e["d_rays_d"], ref["d_rays_d"]), cmp(
            e["grads_fine"]["pts_linears.0.weight"], ref["grads_fine"]["pts_linears.0.weight"])
        print(f"| {name} | {a[0]:.3f} / {a[1]:.4f} | {b[0]:.3f} / {b[1]:.4f} | {c[0]:.3f} / {c[1]:.4f} | "
              f"{float(np.abs(e['rgb_fine'] - ref['rgb_fine']).max()):.1e} |")
    print()
O.BF16_POINTS.update(ALL)
print("""Reading: the backward's own rounding points do not matter -- with the whole backward in fp32 the distance is unchanged.
It comes from the FORWARD operands: rounding only the weights, only the stored activations or only the positional-encoding
features to bf16 each moves the input gradients by 10-19 %.  The input gradient of an L = 10 positional encoding is a sum
of terms scaled by 2^k (up to 512) that largely cancel, so the 2^-9 relative perturbation of any bf16 operand is amplified;
no choice of output precision in the backward (e.g. keeping the three 64-wide encoding-gradient GEMM outputs in fp32)
changes it.  north_star allows bf16/tf32 operands; the remedy would be fp32/tf32 operands throughout, at half the
tensor-core rate.""")
