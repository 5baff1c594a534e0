"""Host-side pose algebra mirror (sparf_b200/camera.py: Lie log maps, Quaternion, to_hom) against vectors produced by the
reference's source/utils/camera.py (tests/golden/make_camera_golden.py)."""
import numpy as np
import torch

from helpers import load_golden
from sparf_b200 import camera as C


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_lie_exp_and_log_maps_match_reference():
    g = load_golden("c9_camera_algebra")
    Rt = C.lie.se3_to_SE3(_t(g["wu"]))
    assert torch.allclose(Rt, _t(g["Rt"]), atol=1e-6)
    assert torch.allclose(C.lie.SO3_to_so3(_t(g["Rt"])[..., :3]), _t(g["so3"]), atol=2e-6)
    assert torch.allclose(C.lie.SE3_to_se3(_t(g["Rt"])), _t(g["se3"]), atol=5e-6)
    # round trip where the log is well conditioned (|w| < pi)
    small = _t(g["wu"])[:, :3].norm(dim=-1) < 3.0
    assert torch.allclose(C.lie.SE3_to_se3(Rt)[small], _t(g["wu"])[small], atol=2e-4)


def test_quaternion_ops_match_reference():
    g = load_golden("c9_camera_algebra")
    q, q2 = _t(g["q"]), _t(g["q2"])
    assert torch.allclose(C.quaternion.q_to_R(q), _t(g["R"]), atol=1e-6)
    assert torch.allclose(C.quaternion.R_to_q(_t(g["R"])), _t(g["q_from_R"]), atol=1e-5)
    assert torch.allclose(C.quaternion.invert(q), _t(g["q_inv"]), atol=1e-6)
    assert torch.allclose(C.quaternion.product(q, q2), _t(g["q_prod"]), atol=1e-6)
    assert torch.equal(C.to_hom(_t(g["wu"])[:, 3:]), _t(g["hom"]))
    # q and R_to_q(q_to_R(q)) describe the same rotation
    qn = torch.nn.functional.normalize(q, dim=-1)
    back = C.quaternion.R_to_q(C.quaternion.q_to_R(q))
    assert torch.allclose(back, qn * torch.sign(qn[:, :1]), atol=1e-5)
