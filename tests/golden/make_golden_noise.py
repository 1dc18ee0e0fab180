"""Golden vectors for the pose-noise initialisation and the pose-error tracking: run the
UNMODIFIED reference `noisy_src/noise.py` (add_noise_to_poses, compute_pose_error) on CPU poses in the authoring
container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_noise.py      ->  tests/golden/noise.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
from noisy_src.noise import NoiseConfig, add_noise_to_poses, compute_pose_error  # noqa: E402

gt = torch.from_numpy(np.load(os.path.join(HERE, "lego_poses.npz"))["ground_truth_poses"])[:40].clone()
gt[7, :3, 3] = 0.0               # a camera at the origin: percentage noise draws nothing for it (std = 0)
out = {"poses": gt.numpy()}
cases = {"rot5_pct5": (5.0, 0.0, 5.0, 42), "rot2_abs": (2.0, 0.1, 0.0, 7), "pct3": (0.0, 0.0, 3.0, 1),
         "rot1p5": (1.5, 0.0, 0.0, 3), "clean": (0.0, 0.0, 0.0, 5)}
for tag, (rot, tabs, pct, seed) in cases.items():
    noisy, infos = add_noise_to_poses(gt, NoiseConfig(rot, tabs, pct, seed=seed))
    info = np.array([[d.get("actual_rotation_deg", 0.0), d.get("actual_translation_norm", 0.0)] for d in infos], np.float64)
    err = np.array([[e["rotation_error_deg"], e["translation_error"]]
                    for e in (compute_pose_error(gt[i], noisy[i]) for i in range(gt.shape[0]))], np.float64)
    out[f"{tag}_cfg"] = np.array([rot, tabs, pct, seed], np.float64)
    out[f"{tag}_noisy"], out[f"{tag}_info"], out[f"{tag}_err"] = noisy.numpy(), info, err
    print(tag, info.mean(0), err.mean(0))
np.savez_compressed(os.path.join(HERE, "noise.npz"), **out)
