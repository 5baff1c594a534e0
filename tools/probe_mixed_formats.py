#!/usr/bin/env python
"""Does one tcgen05.mma.kind::f16 accept a bf16 A operand with an fp16 B operand (a_format != b_format)?
Run in its own process: an illegal instruction poisons the CUDA context."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from sparf_b200 import _lib

L = _lib.lib()
p = lambda t: ctypes.c_void_p(t.data_ptr())
for rows in (64, 128):
    g = torch.Generator().manual_seed(rows)
    G = (torch.randint(-4, 5, (rows, 128), generator=g).float() * 0.5).cuda()       # exact in bf16 and fp16
    X = (torch.randint(-64, 65, (rows, 128), generator=g).float() / 1024).cuda()    # 2^-10 steps: exact in fp16, NOT in bf16 for most values
    D = torch.full((128, 128), -777.0, device="cuda")
    _lib.check(L.sparf_tc_selftest_tn_mixed(p(G), p(X), rows, p(D), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "mixed")
    torch.cuda.synchronize()
    ref = G.double().t() @ X.double()
    err = (D.double() - ref).abs().max().item()
    ref_bf = G.double().t() @ X.bfloat16().double()
    print("rows %d: max |D - exact| = %.3e   (if X had been read as bf16: %.3e)" % (rows, err, (ref_bf - ref).abs().max().item()))
print("mixed f16 x bf16 operands: OK" if err < 1e-6 else "mixed operands: WRONG RESULT")
