"""Pose-parameter modules vs the reference's own modules, imported from /root/reference when it is mounted (the build
container); on the GPU box the reference is absent and the comparison is skipped (the golden-backed tests remain)."""
import os
import sys

import pytest
import torch

import common
from sparf_b200.poses_models import QuaternionsPoseParameters
from sparf_b200.utils.edict import edict

HAVE_REF = os.path.isdir("/root/reference/source")


def _opt(c2w, rel, rot=True, trans=True):
    opt = edict()
    opt.camera = edict(optimize_c2w=c2w, optimize_trans=trans, optimize_rot=rot, optimize_relative_poses=rel,
                       n_first_fixed_poses=1)
    return opt


def _poses(seed, n=4):
    data = common.make_scene(seed, n, 24, 32)
    return common.perturb_poses(data.pose, seed)


@pytest.mark.parametrize("c2w", [False, True])
@pytest.mark.parametrize("rel", [False, True])
def test_quaternion_pose_model_round_trip_and_gradients(c2w, rel):
    w2c = _poses(3)
    m = QuaternionsPoseParameters(_opt(c2w, rel), 4, w2c, torch.device("cpu"))
    assert torch.allclose(m.get_w2c_poses(), w2c, atol=2e-6)
    assert m.rot_embedding.shape == (4 - (1 if rel else 0), 4)
    loss = (m.get_w2c_poses() * torch.linspace(-1, 1, 48).view(4, 3, 4)).sum()
    loss.backward()
    assert m.rot_embedding.grad.abs().sum() > 0 and m.trans_embedding.grad.abs().sum() > 0


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
@pytest.mark.parametrize("c2w,rel,rot,trans", [(False, False, True, True), (True, True, True, True), (True, False, False, True),
                                               (False, True, True, False)])
def test_quaternion_pose_model_matches_reference_module(c2w, rel, rot, trans):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, "golden", "_shims"), "/root/reference"):
        if p not in sys.path:
            sys.path.insert(0, p)
    from source.models.poses_models.quaternion import QuaternionsPoseParameters as Ref
    w2c = _poses(5)
    opt = _opt(c2w, rel, rot, trans)
    ours, ref = QuaternionsPoseParameters(opt, 4, w2c, torch.device("cpu")), Ref(opt, 4, w2c, torch.device("cpu"))
    assert torch.allclose(torch.as_tensor(ours.rot_embedding), torch.as_tensor(ref.rot_embedding), atol=1e-6)
    with torch.no_grad():       # move both off the initial value the same way
        for m in (ours, ref):
            if rot:
                m.rot_embedding += 0.05 * torch.arange(m.rot_embedding.numel()).view_as(m.rot_embedding).float().cos()
            if trans:
                m.trans_embedding += 0.1
    assert torch.allclose(ours.get_w2c_poses(), ref.get_w2c_poses(), atol=1e-6)
    assert torch.allclose(ours.get_c2w_poses(), ref.get_c2w_poses(), atol=1e-6)
