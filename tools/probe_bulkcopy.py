#!/usr/bin/env python
"""Per-SM cp.async.bulk (L2 -> shared) throughput vs copies in flight and copy size (B200)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparf_b200 import _lib

L = _lib.lib()
src = torch.randint(0, 255, (2 << 20,), dtype=torch.uint8, device="cuda")
cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("grid issuer_warps lanes chunk stages  B/clk/SM  (median over CTAs)")
for grid in (148,):
    for nissue in (1, 3):
        for lanes_code in (0, 1, 2, 4, 8):          # 1, 4, 8, 16, 32 lanes
            for chunk in (16384,):
                for stages in (3, 6):
                    if stages % nissue:
                        continue
                    iters = 510
                    for _ in range(2):
                        _lib.check(L.sparf_tc_bulkcopy_probe(ctypes.c_void_p(src.data_ptr()), src.numel(), stages, chunk,
                                                             iters | (nissue << 24) | (lanes_code << 28), grid,
                                                             ctypes.c_void_p(cyc.data_ptr()), st), "probe")
                    torch.cuda.synchronize()
                    c = cyc[:grid].float().median().item()
                    print("%4d %12d %5d %6d %6d  %8.1f" % (grid, nissue, max(1, lanes_code * 4), chunk, stages, iters * chunk / c))
