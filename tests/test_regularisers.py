"""Default-off regularisers: mip-NeRF-360 distortion loss + depth-patch smoothness.
Golden c8_hier_regularisers was produced by the reference's own BasePhotoandReguLoss.compute_loss."""
import pytest
import torch

from helpers import load_golden, rel_err, replay_graph, replay_oracle
from oracle import sparf_oracle as O


def test_oracle_regulariser_terms_match_reference():
    out, loss, grads, gold = replay_oracle("c8_hier_regularisers")
    assert abs(float(out["loss_distortion"]) - float(gold["loss_distortion"])) <= 2e-6 * abs(float(gold["loss_distortion"]))
    assert abs(float(out["loss_depth_patch"]) - float(gold["loss_depth_patch"])) <= 2e-6 * abs(float(gold["loss_depth_patch"]))
    assert abs(float(loss) - float(gold["loss"])) <= 2e-6 * abs(float(gold["loss"]))


@pytest.mark.gpu
@pytest.mark.parametrize("decreasing", [False, True])
@pytest.mark.parametrize("S", [2, 33, 128, 257])
def test_distortion_kernel_matches_quadratic_form(S, decreasing):
    """O(S) prefix-sum kernel vs the reference's [S-1, S-1] pair matrix (restated in the oracle, fp64), values and
    gradients w.r.t. both inputs; decreasing t = inverse-depth sampling."""
    from sparf_b200 import ops
    g = torch.Generator().manual_seed(S)
    t = torch.sort(torch.rand(3, 7, S, 1, generator=g) * 4 + 0.5, dim=2, descending=decreasing).values
    w = torch.rand(3, 7, S, 1, generator=g) ** 3
    t64, w64 = t.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = O.distortion_loss(t64, w64)
    ref.backward()
    tc, wc = t.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    got = ops.distortion_loss(tc, wc)
    (got * 1.7).backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-9
    assert rel_err(wc.grad.cpu().double() / 1.7, w64.grad) <= 2e-5
    assert rel_err(tc.grad.cpu().double() / 1.7, t64.grad) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["simt_fp32", "tc_3x"])
def test_regularised_step_matches_reference(engine):
    out, loss, grads, gold = replay_graph("c8_hier_regularisers", engine=engine)
    assert abs(float(out["loss_distortion"]) - float(gold["loss_distortion"])) <= 1e-3 * abs(float(gold["loss_distortion"]))
    assert abs(float(out["loss_depth_patch"]) - float(gold["loss_depth_patch"])) <= 1e-3 * abs(float(gold["loss_depth_patch"]))
    assert abs(float(loss) - float(gold["loss"])) <= 1e-4 * abs(float(gold["loss"]))
    for k, g in grads.items():
        if k.endswith("bias"):
            assert rel_err(g.cpu(), torch.from_numpy(gold[k])) <= 6e-2, k
