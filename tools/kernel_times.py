#!/usr/bin/env python
"""Per-kernel device time of one MLP training step / inference forward (torch.profiler = CUPTI activity records, no
replay, warm caches).  Usage: [SPARF_KT_ENGINE=tc_3x|tc_3x_w1] [SPARF_KT_NOPOSE=1] python tools/kernel_times.py [R S]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile

import common
from sparf_b200 import _lib, ops


def main():
    R, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1023, 128)
    opt = common.make_opt(S=S)
    sd = common.det_weights(opt, 0)
    keys = sum([["mlp_feat.%d.weight" % i, "mlp_feat.%d.bias" % i] for i in range(8)], []) + \
        ["mlp_rgb.0.weight", "mlp_rgb.0.bias", "mlp_rgb.1.weight", "mlp_rgb.1.bias"]
    params = [sd[k].cuda().requires_grad_(True) for k in keys]
    pose = os.environ.get("SPARF_KT_NOPOSE", "0") != "1"      # ray gradients (camera-pose optimisation) on / off
    o = (torch.randn(R, 3, device="cuda") * 0.3).requires_grad_(pose)
    d = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda"), dim=-1).requires_grad_(pose)
    t = torch.sort(torch.rand(R, S, device="cuda") * 4 + 1.2, dim=1).values
    spec = ops.MLPSpec()
    eng = _lib.ENGINES[os.environ.get("SPARF_KT_ENGINE", "tc_3x")]
    gs, gc = torch.randn(R, S, device="cuda"), torch.randn(R, S, 3, device="cuda")

    def step():
        for p in params:
            p.grad = None
        s, c = ops.mlp_forward(spec, o, d, t, params, engine=eng)
        torch.autograd.backward([s, c], [gs, gc])

    def infer():
        with torch.no_grad():
            ops.mlp_forward(spec, o, d, t, params, engine=eng)

    for name, fn in (("training step (forward with tape + backward)", step), ("inference forward", infer)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n = 5
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
        agg = defaultdict(lambda: [0, 0.0])
        for ev in prof.events():
            if ev.device_type.name == "CUDA":
                agg[ev.name][0] += 1
                agg[ev.name][1] += ev.device_time
        tot = sum(v[1] for v in agg.values()) / n
        print("== %s: %.1f us of kernels per call" % (name, tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("  %8.1f us  x%-3d %s" % (v[1] / n, v[0] // n, k[:110]))


if __name__ == "__main__":
    main()
