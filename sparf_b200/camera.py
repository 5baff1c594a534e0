"""Pose algebra + ray generation: the subset of source/utils/camera.py on the differentiable path.

The [B,3,4] pose chain (inversion, composition, se(3) exponential, 9-D Gram-Schmidt) stays in torch:
it is a handful of 3x4 matrices per step and receives dL/d(pose_w2c) from the ray-generation kernel
(csrc/elementwise.cu: raygen_bwd_kernel).  Ray generation itself runs on the kernel.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from . import ops


class Pose:
    """[...,3,4] = [R|t] operations (camera.py:67-115)."""

    def __call__(self, R=None, t=None):
        assert R is not None or t is not None
        if R is None:
            t = torch.as_tensor(t)
            R = torch.eye(3, device=t.device).repeat(*t.shape[:-1], 1, 1)
        elif t is None:
            R = torch.as_tensor(R)
            t = torch.zeros(R.shape[:-1], device=R.device)
        else:
            R, t = torch.as_tensor(R), torch.as_tensor(t)
        assert R.shape[:-1] == t.shape and R.shape[-2:] == (3, 3)
        return torch.cat([R.float(), t.float()[..., None]], dim=-1)

    def invert(self, pose: torch.Tensor, use_inverse: bool = False) -> torch.Tensor:
        R, t = pose[..., :3], pose[..., 3:]
        R_inv = R.inverse() if use_inverse else R.transpose(-1, -2)
        return self(R=R_inv, t=(-R_inv @ t)[..., 0])

    def compose(self, pose_list: List[torch.Tensor]) -> torch.Tensor:
        """pose_new(x) = pose_N o ... o pose_1(x)."""
        out = pose_list[0]
        for p in pose_list[1:]:
            out = self.compose_pair_b_at_a(out, p)
        return out

    def compose_pair_b_at_a(self, pose_a, pose_b):
        R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
        R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
        return self(R=R_b @ R_a, t=(R_b @ t_a + t_b)[..., 0])


class Lie:
    """so(3)/se(3) exponential with the reference's 10-term Taylor series (camera.py:117-205)."""

    @staticmethod
    def _series(x, start_factor, nth=10):
        # sum_i (-1)^i x^(2i) / d_i with d_i built from consecutive integer pairs
        ans = torch.zeros_like(x)
        denom = 1.0
        for i in range(nth + 1):
            a, b = start_factor(i)
            if a is not None:
                denom *= a * b
            ans = ans + (-1) ** i * x ** (2 * i) / denom
        return ans

    def taylor_A(self, x, nth=10):  # sin(x)/x
        return self._series(x, lambda i: (2 * i, 2 * i + 1) if i > 0 else (None, None), nth)

    def taylor_B(self, x, nth=10):  # (1-cos x)/x^2
        return self._series(x, lambda i: (2 * i + 1, 2 * i + 2), nth)

    def taylor_C(self, x, nth=10):  # (x-sin x)/x^3
        return self._series(x, lambda i: (2 * i + 2, 2 * i + 3), nth)

    def skew_symmetric(self, w):
        w0, w1, w2 = w.unbind(dim=-1)
        O = torch.zeros_like(w0)
        return torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1),
                            torch.stack([-w1, w0, O], dim=-1)], dim=-2)

    def so3_to_SO3(self, w):
        wx = self.skew_symmetric(w)
        theta = w.norm(dim=-1)[..., None, None]
        I = torch.eye(3, device=w.device, dtype=torch.float32)
        return I + self.taylor_A(theta) * wx + self.taylor_B(theta) * wx @ wx

    def se3_to_SE3(self, wu):
        w, u = wu.split([3, 3], dim=-1)
        wx = self.skew_symmetric(w)
        theta = w.norm(dim=-1)[..., None, None]
        I = torch.eye(3, device=w.device, dtype=torch.float32)
        A, B, C = self.taylor_A(theta), self.taylor_B(theta), self.taylor_C(theta)
        R = I + A * wx + B * wx @ wx
        V = I + B * wx + C * wx @ wx
        return torch.cat([R, V @ u[..., None]], dim=-1)

    def SO3_to_so3(self, R, eps=1e-7):
        """Rotation matrix -> axis-angle vector (camera.py:133-140): theta from the clamped trace (taken modulo pi, the
        reference's guard against theta == pi), w = vee((R - R^T) / (2 sin(theta)/theta))."""
        trace = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
        theta = torch.remainder(torch.acos(((trace - 1) / 2).clamp(-1 + eps, 1 - eps)), math.pi)[..., None, None]
        lnR = 1 / (2 * self.taylor_A(theta) + 1e-8) * (R - R.transpose(-2, -1))
        return torch.stack([lnR[..., 2, 1], lnR[..., 0, 2], lnR[..., 1, 0]], dim=-1)

    def SE3_to_se3(self, Rt, eps=1e-8):
        """[R|t] -> (w, u) with t = V(w) u (camera.py:159-170)."""
        R, t = Rt.split([3, 1], dim=-1)
        w = self.SO3_to_so3(R)
        wx = self.skew_symmetric(w)
        theta = w.norm(dim=-1)[..., None, None]
        I = torch.eye(3, device=w.device, dtype=torch.float32)
        A, B = self.taylor_A(theta), self.taylor_B(theta)
        invV = I - 0.5 * wx + (1 - A / (2 * B)) / (theta ** 2 + eps) * wx @ wx
        return torch.cat([w, (invV @ t)[..., 0]], dim=-1)


class Quaternion:
    """Unit-quaternion pose parametrisation, scalar first (camera.py:207-291)."""

    def q_to_R(self, q):
        qa, qb, qc, qd = torch.nn.functional.normalize(q, dim=-1).unbind(dim=-1)
        rows = [[1 - 2 * (qc ** 2 + qd ** 2), 2 * (qb * qc - qa * qd), 2 * (qa * qc + qb * qd)],
                [2 * (qb * qc + qa * qd), 1 - 2 * (qb ** 2 + qd ** 2), 2 * (qc * qd - qa * qb)],
                [2 * (qb * qd - qa * qc), 2 * (qa * qb + qc * qd), 1 - 2 * (qb ** 2 + qc ** 2)]]
        return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

    def R_to_q(self, R, eps=1e-8):
        """Batched [B,3,3] rotations -> quaternions with w >= 0 via the dominant eigenvector of the symmetric 4x4
        matrix built from R (host-side numpy, not differentiable, like the reference: camera.py:242-275)."""
        import numpy as np
        dev = R.device if torch.is_tensor(R) else None
        Rn = R.detach().cpu().numpy() if torch.is_tensor(R) else np.asarray(R)
        assert Rn.ndim == 3, "R_to_q expects a batch of rotation matrices"
        out = []
        for M in Rn:
            (xx, yx, zx), (xy, yy, zy), (xz, yz, zz) = M        # row-major unpack, i.e. R.flat order
            K = np.array([[xx - yy - zz, 0, 0, 0],
                          [yx + xy, yy - xx - zz, 0, 0],
                          [zx + xz, zy + yz, zz - xx - yy, 0],
                          [yz - zy, zx - xz, xy - yx, xx + yy + zz]]) / 3.0
            vals, vecs = np.linalg.eigh(K)
            q = vecs[[3, 0, 1, 2], np.argmax(vals)]
            out.append(-q if q[0] < 0 else q)
        q = np.stack(out, axis=0)
        return torch.from_numpy(q).to(dev).float() if dev is not None else q

    def invert(self, q):
        qa, qb, qc, qd = q.unbind(dim=-1)
        return torch.stack([qa, -qb, -qc, -qd], dim=-1) / q.norm(dim=-1, keepdim=True) ** 2

    def product(self, q1, q2):
        a1, b1, c1, d1 = q1.unbind(dim=-1)
        a2, b2, c2, d2 = q2.unbind(dim=-1)
        return torch.stack([a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
                            a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                            a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
                            a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2], dim=-1)


def to_hom(X):
    """Homogeneous coordinates (camera.py:297-300)."""
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)


pose = Pose()
lie = Lie()
quaternion = Quaternion()


def get_center_and_ray(pose_w2c: torch.Tensor, H: int, W: int, intr: torch.Tensor, ray_idx=None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """camera.get_center_and_ray (camera.py:347-382) evaluated only at `ray_idx` (all H*W pixels if None)."""
    if ray_idx is None:
        ray_idx = torch.arange(H * W, device=pose_w2c.device)
    return ops.raygen(pose_w2c, intr, W, ray_idx=ray_idx)


def get_center_and_ray_at_pixels(pose_w2c, pixels, intr):
    """camera.get_center_and_ray_at_pixels (camera.py:384-416): float pixels, no +0.5."""
    return ops.raygen(pose_w2c, intr, 0, pixels=pixels)


def get_3D_points_from_depth(center, ray, depth, multi_samples: bool = False):
    """x = c + d*v (camera.py:418-437)."""
    if multi_samples:
        center, ray = center[:, :, None], ray[:, :, None]
    return center + ray * depth
