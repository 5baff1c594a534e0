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
print("grid chunk stages  B/clk/SM  (median over CTAs)")
for grid in (1, 148):
    for chunk in (4096, 8192, 16384, 32768):
        for stages in (1, 2, 3, 4, 6, 8, 12):
            if stages * chunk > 192 * 1024:
                continue
            iters = 512
            for _ in range(2):
                _lib.check(L.sparf_tc_bulkcopy_probe(ctypes.c_void_p(src.data_ptr()), src.numel(), stages, chunk, iters, grid,
                                                     ctypes.c_void_p(cyc.data_ptr()), st), "probe")
            torch.cuda.synchronize()
            c = cyc[:grid].float().median().item()
            print("%4d %6d %6d  %8.1f" % (grid, chunk, stages, iters * chunk / c))
