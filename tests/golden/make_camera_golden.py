#!/usr/bin/env python
"""Golden vectors for the host-side pose algebra (Lie log maps, quaternions) from the reference's own
source/utils/camera.py.  Run in the build container (needs /root/reference):  python make_camera_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, "/root/reference")
from source.utils import camera as C  # noqa: E402


def main():
    g = torch.Generator().manual_seed(11)
    w = torch.randn(16, 3, generator=g) * torch.tensor([0.02, 0.5, 1.5])[torch.randint(0, 3, (16, 1), generator=g)]
    u = torch.randn(16, 3, generator=g)
    wu = torch.cat([w, u], dim=-1)
    Rt = C.lie.se3_to_SE3(wu)
    q = torch.randn(16, 4, generator=g)
    q2 = torch.randn(16, 4, generator=g)
    R = C.quaternion.q_to_R(q)
    out = dict(wu=wu, Rt=Rt, so3=C.lie.SO3_to_so3(Rt[..., :3]), se3=C.lie.SE3_to_se3(Rt), q=q, q2=q2, R=R,
               q_from_R=C.quaternion.R_to_q(R), q_inv=C.quaternion.invert(q), q_prod=C.quaternion.product(q, q2),
               hom=C.to_hom(u))
    np.savez_compressed(os.path.join(HERE, "c9_camera_algebra.npz"), **{k: v.numpy() for k, v in out.items()})
    print({k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
