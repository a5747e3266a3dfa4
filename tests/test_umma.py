"""tcgen05 building block (mvin_b200/csrc/umma.cuh): the 3xTF32 tensor-core product against an fp64 reference.
Tolerance: 2e-6 of the largest |C| entry (three-product split keeps ~21 mantissa bits; plain TF32 would be ~5e-4)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dw_lane(i: int, D: int) -> int:
    """Accumulator lane of output row i of a tcgen05 MMA with M = 64 (cta_group::1): 16 rows per 32-lane quadrant
    (measured by test_umma_dw_probe; see csrc/umma.cuh)."""
    return 32 * (i // 16) + i % 16


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


@pytest.mark.parametrize("D,M", [(64, 128), (64, 700), (32, 128), (32, 300)])
def test_umma_dw_probe(D, M):
    """Transposed product dW = A^T G (single-product TF32 probe, M = 64): K-major transposed staging works and row i
    of the 64-row accumulator sits in lane 32 (i / 16) + i % 16; the MN-major view of the K-major tiles (variant 3)
    is NOT usable for kind::tf32 without swizzle -- the tensor core returns zeros -- which is why the weight
    gradients stay on the mma.sync path (level.cuh)."""
    from mvin_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(7 * D + M)
    A = torch.randn(M, D, generator=g)
    G = torch.randn(M, D, generator=g)
    dA, dG = A.cuda(), G.cuda()
    ref = (A.double().t() @ G.double()).numpy()
    scale = np.abs(ref).max()
    dump = torch.full((128, D), float("nan"), device="cuda")
    rc = lib.mvin_test_umma_dw(ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dG.data_ptr()),
                               ctypes.c_void_p(dump.data_ptr()), ctypes.c_int64(M), ctypes.c_int32(D), 0, None)
    assert rc == 0, lib.mvin_last_error()
    torch.cuda.synchronize()
    got = dump.cpu().numpy()
    for i in range(D):
        assert np.abs(got[dw_lane(i, D)] - ref[i]).max() / scale < 5e-3, i      # one TF32 product: ~1e-3


@pytest.mark.parametrize("D,M,NP", [(64, 128, 2), (64, 1000, 2), (64, 700, 3), (32, 333, 2), (32, 128, 3)])
def test_umma_bf16_split_serves_both_contractions(D, M, NP):
    """Split-bf16 operands (csrc/umma_bf.cuh): ONE staged tile per operand is the K-major operand of C = G . W^T and the
    MN-major operand of dW = A^T . G (accumulated over the 128-row tiles in tensor memory).  NP = 2 planes keep 16
    mantissa bits (error of order 1e-5 of the largest entry), NP = 3 all 24."""
    from mvin_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(11 * D + M + NP)
    A = torch.randn(M, D, generator=g)
    G = torch.randn(M, D, generator=g) * (torch.rand(M, D, generator=g) > 0.4)          # ReLU-masked gradient
    W = torch.randn(D, D, generator=g) / D ** 0.5
    dA, dG, dW_ = A.cuda(), G.cuda(), W.cuda()
    dC = torch.full((M, D), float("nan"), device="cuda")
    dump = torch.full((128, D), float("nan"), device="cuda")
    vp = ctypes.c_void_p
    rc = lib.mvin_test_umma_bf16(vp(dA.data_ptr()), vp(dG.data_ptr()), vp(dW_.data_ptr()), vp(dC.data_ptr()),
                                 vp(dump.data_ptr()), ctypes.c_int64(M), ctypes.c_int32(D), ctypes.c_int32(NP), None)
    assert rc == 0, lib.mvin_last_error()
    torch.cuda.synchronize()
    tol = 3e-5 if NP == 2 else 5e-6       # NP = 3 measured on B200: 3.0e-6 at M = 700 (the dropped lo x lo cross terms)
    ref_c = (G.double() @ W.double().t()).numpy()
    assert np.abs(dC.cpu().numpy() - ref_c).max() / np.abs(ref_c).max() < tol
    ref_w = (A.double().t() @ G.double()).numpy()
    got = dump.cpu().numpy()
    for i in range(D):
        assert np.abs(got[dw_lane(i, D)] - ref_w[i]).max() / np.abs(ref_w).max() < tol, i
