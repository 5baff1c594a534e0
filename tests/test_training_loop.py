"""End-to-end: a few hundred complete training iterations (device-side ray sampling -> render -> loss module -> backward ->
fused clip + Adam + schedule), each replayed as ONE CUDA graph, must actually fit a synthetic scene
(tools/train_synthetic.py: a teacher NeRF renders the views, a fresh student learns them)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("fine", [0, 1])
def test_training_loop_converges(fine):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_synthetic
    first, last = train_synthetic.main(["--steps", "300", "--quiet", "--fine", str(fine), "--rays", "768"])
    print("fine=%d: loss %.5f -> %.5f" % (fine, first, last))
    assert last == last and first == first            # finite
    assert last < 0.6 * first, (first, last)
