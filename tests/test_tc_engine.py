"""tcgen05 engine: hardware self-test of the Blackwell primitives, then the fused MLP kernels against
the SIMT fp32 engine (same C ABI, same inputs, on the device) and the reference goldens."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("K", [64, 128, 256])
def test_tcgen05_selftest_gemm_exact(K):
    """Small-integer operands are exact in bf16 and their dot products exact in fp32: bit-exact result
    proves operand layout, descriptors, K stepping, bulk copy and TMEM addressing."""
    from sparf_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(K)
    A = torch.randint(-4, 5, (128, K), generator=g).float().cuda()
    B = torch.randint(-4, 5, (128, K), generator=g).float().cuda()
    packed = torch.zeros(128 * K * 2, dtype=torch.uint8, device="cuda")
    D = torch.full((128, 128), -777.0, device="cuda")
    _lib.check(L.sparf_tc_selftest(_p(A), _p(B), K, _p(packed), _p(D), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "tc_selftest")
    torch.cuda.synchronize()
    ref = A @ B.t()
    assert torch.equal(D, ref), (D - ref).abs().max().item()
