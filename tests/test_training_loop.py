"""End-to-end: a few hundred complete training iterations (device-side ray sampling -> render -> loss module -> backward ->
fused clip + Adam + schedule), each replayed as ONE CUDA graph, must actually fit a synthetic scene
(tools/train_synthetic.py: a teacher NeRF renders the views, a fresh student learns them)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("fine,poses,engine", [(0, 0, "auto"), (1, 0, "auto"), (0, 1, "auto"), (0, 0, "tc_3x_w1"), (0, 1, "tc_3x_w1")])
def test_training_loop_converges(fine, poses, engine):
    """coarse only / hierarchical / joint pose-NeRF (BARF mask advancing on the device, second fused-Adam group for the
    9-D pose embeddings); also with the non-default reduced-precision weight-gradient engine (same loss curve to ~1 %)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sparf_b200
    import train_synthetic
    try:
        res = train_synthetic.main(["--steps", "300", "--quiet", "--fine", str(fine), "--poses", str(poses), "--rays", "768",
                                    "--engine", engine])
    finally:
        sparf_b200.set_engine("auto")
    first, last = res[0], res[1]
    print("fine=%d poses=%d engine=%s: loss %.5f -> %.5f" % (fine, poses, engine, first, last))
    assert last == last and first == first            # finite
    assert last < (0.85 if poses else 0.6) * first, (first, last)
