"""Cold-L2 repetitions of the taped MLP training step (tools/stress_chain.py): every repetition must reproduce the
first one bit-exactly in the forward outputs.  This is the test that exposes mbarrier protocol races between the warp
roles of the chain kernels (a lapped waiter hung one step in ~50 before the ring barriers counted every waiter)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cold_l2_repetitions_are_reproducible(monkeypatch, capsys):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stress_chain
    monkeypatch.setattr(sys, "argv", ["stress_chain.py", "40"])
    stress_chain.main()
    assert "stress ok: 40 repetitions" in capsys.readouterr().out
