"""SE(3) axis-angle pose-refinement parameters behind the reference's interface
(noisy_src/train_pose_opt.py:53-271 CameraPoseParameters)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


class CameraPoseParameters(nn.Module):
    """R = exp(omega) @ R_init, t = t_init + delta_t.  Same attribute and state_dict names as the
    reference: `rotation_deltas (N,3)`, `translation_deltas (N,3)` (Parameters, or buffers when not
    learned), buffer `initial_poses (N,4,4)`; all deltas zero-initialised."""

    def __init__(self, initial_poses: torch.Tensor, learn_rotation: bool = True, learn_translation: bool = True):
        super().__init__()
        self.n_poses = initial_poses.shape[0]
        self.learn_rotation = learn_rotation
        self.learn_translation = learn_translation
        self.register_buffer("initial_poses", initial_poses.clone())
        zeros = lambda: torch.zeros(self.n_poses, 3, device=initial_poses.device)
        if learn_rotation:
            self.rotation_deltas = nn.Parameter(zeros())
        else:
            self.register_buffer("rotation_deltas", zeros())
        if learn_translation:
            self.translation_deltas = nn.Parameter(zeros())
        else:
            self.register_buffer("translation_deltas", zeros())

    def axis_angle_to_rotation_matrix(self, axis_angle: torch.Tensor) -> torch.Tensor:
        """Rodrigues with the theta < 1e-6 -> identity (zero-gradient) branch (train_pose_opt.py:122-163)."""
        shp = axis_angle.shape[:-1]
        w = axis_angle.reshape(-1, 3)
        eye = torch.eye(4, device=w.device).expand(w.shape[0], 4, 4).contiguous()
        P = ops.SE3Poses.apply(eye, w, torch.zeros_like(w), None, True, False)
        return P[:, :3, :3].reshape(*shp, 3, 3)

    def get_poses(self, indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(N,4,4) or (len(indices),4,4) current camera-to-world poses (train_pose_opt.py:186-226)."""
        return ops.SE3Poses.apply(self.initial_poses, self.rotation_deltas, self.translation_deltas, indices,
                                  self.learn_rotation, self.learn_translation)

    def get_all_poses(self) -> torch.Tensor:
        return self.get_poses()

    @torch.no_grad()
    def compute_pose_errors(self, ground_truth_poses: torch.Tensor, indices: Optional[torch.Tensor] = None
                            ) -> Dict[str, float]:
        """Geodesic rotation error (deg) and translation L2 statistics (train_pose_opt.py:232-271,
        noise.py:237-268): one launch over all poses (`rn_pose_errors`) and one host copy instead of 200
        `.item()` synchronisations; the statistics are numpy float64 over the per-pose values, as in the reference."""
        from .noise import compute_pose_errors_batch
        cur = self.get_poses(indices)
        gt = ground_truth_poses.to(cur.device)
        if indices is not None:
            gt = gt[indices]
        err = compute_pose_errors_batch(gt, cur).cpu().numpy().astype(np.float64)
        rot, tra = err[:, 0], err[:, 1]
        return {"rotation_error_mean": float(np.mean(rot)), "rotation_error_std": float(np.std(rot)),
                "rotation_error_max": float(np.max(rot)), "translation_error_mean": float(np.mean(tra)),
                "translation_error_std": float(np.std(tra)), "translation_error_max": float(np.max(tra))}
