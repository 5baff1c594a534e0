"""9-D pose embedding (translation + first two rows of R, Gram-Schmidt), the reference's default
learnable pose model: source/models/poses_models/two_columns.py.  Same class name (including the
reference's spelling), constructor, attributes (`pose_embedding` / `trans_embedding` / `rot_embedding`)
and methods, so checkpoints and the pose optimiser are interchangeable."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import camera


def pose_to_d9(pose: torch.Tensor) -> torch.Tensor:
    """[N,3,4] -> [N,9] = (t, R[0,:], R[1,:])   (two_columns.py:23-39)."""
    return torch.cat((pose[:, :3, -1], pose[:, :2, :3].reshape(pose.shape[0], -1)), -1)


def r6d2mat(d6: torch.Tensor) -> torch.Tensor:
    """Two rows -> rotation matrix by Gram-Schmidt, third row = cross product (two_columns.py:42-62)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


class FirstTwoColunmnsPoseParameters(nn.Module):
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

    def init_poses_embed(self):
        init = self.initial_poses_c2w if self.optimize_c2w else self.initial_poses_w2c
        embed = pose_to_d9(init[self._n_fixed():])
        if self.optimize_rot and self.optimize_trans:
            self.pose_embedding = nn.Parameter(embed)
        elif self.optimize_rot:
            self.trans_embedding = embed[:, :3]
            self.rot_embedding = nn.Parameter(embed[:, 3:])
        elif self.optimize_trans:
            self.trans_embedding = nn.Parameter(embed[:, :3])
            self.rot_embedding = embed[:, 3:]
        else:
            raise ValueError("Either the trans or the rot must be optimized")

    def get_initial_w2c(self):
        return self.initial_poses_w2c

    def _embedded_poses(self, fixed_init):
        if self.optimize_rot and self.optimize_trans:
            t, r = self.pose_embedding[:, :3], self.pose_embedding[:, 3:]
        else:
            t, r = self.trans_embedding, self.rot_embedding
        poses = torch.cat((r6d2mat(r)[:, :3, :3], t[..., None]), -1)
        n_fixed = self._n_fixed()
        if n_fixed > 0 or self.opt.camera.optimize_relative_poses:
            poses = torch.cat((fixed_init[:n_fixed], poses), dim=0)
            assert poses.shape[0] == self.nbr_poses
        return poses

    def get_c2w_poses(self) -> torch.Tensor:
        if self.optimize_c2w:
            return self._embedded_poses(self.initial_poses_c2w)
        return camera.pose.invert(self.get_w2c_poses())

    def get_w2c_poses(self) -> torch.Tensor:
        if not self.optimize_c2w:
            return self._embedded_poses(self.initial_poses_w2c)
        return camera.pose.invert(self.get_c2w_poses())
