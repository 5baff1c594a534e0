"""Checkpoint compatibility: our Graph's state_dict has exactly the reference Graph's keys and shapes
(fixture produced by tests/golden/make_state_dict_keys.py from the reference), and a save -> load round trip through
torch.save preserves every tensor."""
import io
import json
import os

import pytest
import torch

import common

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("tag,kw", [("coarse_only", dict(fine=False)), ("hierarchical_c2f", dict(fine=True, barf_c2f=(0.1, 0.5)))])
def test_state_dict_keys_and_round_trip(tag, kw):
    from sparf_b200.renderer import Graph
    ref = json.load(open(os.path.join(HERE, "golden", "state_dict_keys.json")))[tag]
    opt = common.make_opt(S=64, **kw)
    net = Graph(opt, torch.device("cpu"))
    sd = net.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == ref
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    net2 = Graph(opt, torch.device("cpu"))
    net2.load_state_dict(torch.load(buf))
    for k, v in net2.state_dict().items():
        assert torch.equal(v, sd[k]), k
