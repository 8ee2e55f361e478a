"""The tcgen05 (3xTF32) contraction core and the tensor-core staged solver against float64 / the other kernel families."""
import numpy as np
import pytest
import torch

import odecol

pytestmark = pytest.mark.gpu
DEV = "cuda"


# the last three: more than 12 population tiles, so the grouped tile order (tile_coords in csrc/stage_tc.cuh) is walked with
# a ragged last group (14 = 12 + 2, 25 = 12 + 12 + 1, 13 = 12 + 1 with chunked accumulation, K > 768)
@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (128, 128, 64), (512, 300, 577), (200, 1000, 130), (1024, 2368, 608),
                                   (1792, 300, 96), (3100, 1200, 64), (1664, 240, 800)])
def test_tc_contract_matches_float64(M, N, K):
    ext = odecol._native.ext()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * torch.rand(M, K, generator=g) * 30).to(DEV)
    B = (torch.randn(N, K, generator=g).abs() * 5).to(DEV)
    C = ext.tc_contract(A, B)
    torch.cuda.synchronize()
    ref = B.double() @ A.double().T
    mag = B.double().abs() @ A.double().abs().T                     # sum_k |a||b|: the natural error scale of a dot product
    err = float(((C.double() - ref).abs() / mag).max())
    fp32 = float((((B @ A.T).double() - ref).abs() / mag).max())
    print(f"\ntc_contract {M}x{N}x{K}: max err / sum|a||b| = {err:.2e} (cuBLAS fp32: {fp32:.2e})")
    assert err < 2e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (512, 577, 1000), (200, 96, 4096)])
def test_tc_contract_tn_matches_float64(M, N, K):
    """The MN-major variant (both operands read along their contiguous dimension) used for dW_aug."""
    ext = odecol._native.ext()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(K, M, generator=g) * 3).to(DEV)
    B = (torch.randn(K, N, generator=g).abs() * 5).to(DEV)
    C = ext.tc_contract_tn(A, B)
    torch.cuda.synchronize()
    ref = A.double().T @ B.double()
    mag = A.double().abs().T @ B.double().abs()
    err = float(((C.double() - ref).abs() / mag).max())
    print(f"\ntc_contract_tn {M}x{N}x{K}: max err / sum|a||b| = {err:.2e}; |C|max {float(C.abs().max()):.3e} |ref|max {float(ref.abs().max()):.3e}")
    assert err < 2e-6
