#!/usr/bin/env python
"""Device timing of the MLP entry points (CUDA events, warm): forward per engine, taped forward, backward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch

import common
from sparf_b200 import _lib, ops


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    R, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1023, 128)
    opt = common.make_opt(S=S)
    sd = common.det_weights(opt, 0)
    keys = sum([["mlp_feat.%d.weight" % i, "mlp_feat.%d.bias" % i] for i in range(8)], []) + \
        ["mlp_rgb.0.weight", "mlp_rgb.0.bias", "mlp_rgb.1.weight", "mlp_rgb.1.bias"]
    params = [sd[k].cuda().requires_grad_(True) for k in keys]
    o = torch.randn(R, 3, device="cuda") * 0.3
    d = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda"), dim=-1)
    t = torch.sort(torch.rand(R, S, device="cuda") * 4 + 1.2, dim=1).values
    spec = ops.MLPSpec()
    gs, gc = torch.randn(R, S, device="cuda"), torch.randn(R, S, 3, device="cuda")
    flop_fwd = R * S * 2 * 527872
    for name, eng in (("tc_3x", _lib.ENGINE_TC_3X), ("tc_1x", _lib.ENGINE_TC_1X), ("simt", _lib.ENGINE_SIMT_FP32)):
        with torch.no_grad():
            ms = timeit(lambda: ops.mlp_forward(spec, o, d, t, params, engine=eng), n=20 if eng != _lib.ENGINE_SIMT_FP32 else 3)
        print("forward  %-6s %8.3f ms   %7.1f TFLOP/s algorithmic" % (name, ms, flop_fwd / ms / 1e9))
    for tape in (True, False):
        ops.USE_TAPE[0] = tape

        def step():
            for p in params:
                p.grad = None
            s, c = ops.mlp_forward(spec, o, d, t, params, engine=_lib.ENGINE_TC_3X)
            torch.autograd.backward([s, c], [gs, gc])
        ms = timeit(step, n=10)
        print("fwd+bwd  tc_3x tape=%-5s %8.3f ms   %7.1f TFLOP/s algorithmic" % (tape, ms, 3 * flop_fwd / ms / 1e9))


if __name__ == "__main__":
    main()
