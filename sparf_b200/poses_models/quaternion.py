"""Unit-quaternion pose parametrisation: source/models/poses_models/quaternion.py.

`rot_embedding` [N,4] (w, x, y, z) and `trans_embedding` [N,3] hold the optimised frame's rotation and translation
(camera-to-world if `opt.camera.optimize_c2w`, else world-to-camera); either can be frozen
(`optimize_rot` / `optimize_trans`), and with `optimize_relative_poses` the first `n_first_fixed_poses` poses keep their
initial value and have no embedding.  The quaternion is re-normalised on every read; gradients reach both embeddings
through `q_to_R`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import camera


class QuaternionsPoseParameters(nn.Module):
    def __init__(self, opt, nbr_poses: int, initial_poses_w2c: torch.Tensor, device):
        super().__init__()
        self.opt = opt
        self.optimize_c2w = opt.camera.optimize_c2w
        self.optimize_trans = opt.camera.optimize_trans
        self.optimize_rot = opt.camera.optimize_rot
        self.nbr_poses = nbr_poses
        self.device = device
        self.initial_poses_w2c = initial_poses_w2c
        self.initial_poses_c2w = camera.pose.invert(initial_poses_w2c)
        self.init_poses_embed()

    def _n_fixed(self):
        return self.opt.camera.n_first_fixed_poses if self.opt.camera.optimize_relative_poses else 0

    def _frame_poses(self):
        return self.initial_poses_c2w if self.optimize_c2w else self.initial_poses_w2c

    def init_poses_embed(self):
        poses = self._frame_poses()[self._n_fixed():]
        q = camera.quaternion.R_to_q(poses[:, :3, :3])
        t = poses[:, :3, -1]
        self.rot_embedding = nn.Parameter(q) if self.optimize_rot else q
        self.trans_embedding = nn.Parameter(t) if self.optimize_trans else t

    def _embedded_poses(self):
        """Poses of the optimised frame (c2w or w2c), fixed ones first."""
        q = torch.nn.functional.normalize(self.rot_embedding, dim=-1)
        R = camera.quaternion.q_to_R(q)[:, :3, :3]
        poses = torch.cat((R, self.trans_embedding[..., None]), dim=-1)
        n_fixed = self._n_fixed()
        if n_fixed:
            poses = torch.cat((self._frame_poses()[:n_fixed], poses), dim=0)
            assert poses.shape[0] == self.nbr_poses
        return poses

    def get_c2w_poses(self) -> torch.Tensor:
        return self._embedded_poses() if self.optimize_c2w else camera.pose.invert(self.get_w2c_poses())

    def get_w2c_poses(self) -> torch.Tensor:
        return camera.pose.invert(self._embedded_poses()) if self.optimize_c2w else self._embedded_poses()
