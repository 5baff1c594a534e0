"""se(3) residual composed with the initial pose (BARF-style): source/models/poses_models/axis_rotation.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import camera


class AxisRotationPoseParameters(nn.Module):
    def __init__(self, opt, nbr_poses: int, initial_poses_w2c: torch.Tensor, device):
        super().__init__()
        self.opt = opt
        self.nbr_poses = nbr_poses
        self.device = device
        self.initial_poses_w2c = initial_poses_w2c
        self.init_poses_embed()

    def _n_fixed(self):
        return self.opt.camera.n_first_fixed_poses if self.opt.camera.optimize_relative_poses else 0

    def init_poses_embed(self):
        self.pose_embedding = nn.Parameter(torch.zeros(self.nbr_poses - self._n_fixed(), 6, device=self.device))

    def get_w2c_poses(self) -> torch.Tensor:
        refine = camera.lie.se3_to_SE3(self.pose_embedding)
        n_fixed = self._n_fixed()
        poses = camera.pose.compose([refine, self.initial_poses_w2c[n_fixed:]])
        if self.opt.camera.optimize_relative_poses:
            poses = torch.cat((self.initial_poses_w2c[:n_fixed], poses), dim=0)
            assert poses.shape[0] == self.nbr_poses
        return poses

    def get_c2w_poses(self) -> torch.Tensor:
        return camera.pose.invert(self.get_w2c_poses())
