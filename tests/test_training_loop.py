"""End-to-end: a few hundred complete training iterations (device-side ray sampling -> render -> loss module -> backward ->
fused clip + Adam + schedule), each replayed as ONE CUDA graph, must actually fit a synthetic scene
(tools/train_synthetic.py: a teacher NeRF renders the views, a fresh student learns them)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("fine,poses", [(0, 0), (1, 0), (0, 1)])
def test_training_loop_converges(fine, poses):
    """coarse only / hierarchical / joint pose-NeRF (BARF mask advancing on the device, second fused-Adam group for the
    9-D pose embeddings)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_synthetic
    res = train_synthetic.main(["--steps", "300", "--quiet", "--fine", str(fine), "--poses", str(poses), "--rays", "768"])
    first, last = res[0], res[1]
    print("fine=%d poses=%d: loss %.5f -> %.5f" % (fine, poses, first, last))
    assert last == last and first == first            # finite
    assert last < (0.85 if poses else 0.6) * first, (first, last)
