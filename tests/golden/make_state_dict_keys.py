#!/usr/bin/env python
"""Key / shape listing of the reference Graph's state_dict (checkpoint compatibility fixture).
Run in the build container (needs /root/reference):  python make_state_dict_keys.py"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)
import common  # noqa: E402
from source.models.renderer import Graph  # noqa: E402


def main():
    out = {}
    for tag, kw in (("coarse_only", dict(fine=False)), ("hierarchical_c2f", dict(fine=True, barf_c2f=(0.1, 0.5)))):
        opt = common.make_opt(S=64, **kw)
        net = Graph(opt, torch.device("cpu"))
        out[tag] = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump(out, open(os.path.join(HERE, "state_dict_keys.json"), "w"), indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
