"""tcgen05 building block (mvin_b200/csrc/umma.cuh): the 3xTF32 tensor-core product against an fp64 reference.
Tolerance: 2e-6 of the largest |C| entry (three-product split keeps ~21 mantissa bits; plain TF32 would be ~5e-4)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,M", [(32, 128), (64, 128), (64, 1000), (32, 333)])
def test_umma_gemm_matches_fp64(D, M):
    from mvin_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(D + M)
    A = torch.randn(M, D, generator=g)
    W = torch.randn(D, D, generator=g)
    dA, dW = A.cuda(), W.cuda()
    dC = torch.full((M, D), float("nan"), device="cuda")
    rc = lib.mvin_test_umma_gemm(ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dW.data_ptr()),
                                 ctypes.c_void_p(dC.data_ptr()), ctypes.c_int64(M), ctypes.c_int32(D), None)
    assert rc == 0, lib.mvin_last_error()
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).numpy()
    got = dC.cpu().numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err
