"""The tensor-core contraction of the wide-network path (csrc/gemm_tc.cuh: tcgen05.mma kind::tf32, 3-pass split,
TMEM accumulators, TMA operand loads) against an fp64 product.  The bar is fp32-GEMM accuracy: the split must
recover what a single tf32 pass (10-bit mantissa, error ~5e-4) loses, otherwise north_star's rtol 1e-5 on the
ODE results is out of reach and the path would have to fall back to FFMA."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(eno, A, B, bn, splits=1):
    L = eno._cabi.lib()
    h = eno._cabi.ctx(A.device.index)
    M, K = A.shape
    N = B.shape[0]
    out = torch.full((M, N), float("nan"), device=A.device)
    st = C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)
    rc = L.eno_gemm_tf32x3(h, M, N, K, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0),
                           C.c_void_p(out.data_ptr()), out.stride(0), bn, splits, st)
    eno._cabi.check(rc, h, "eno_gemm_tf32x3")
    return out


# (M, N, K): one exact tile; ragged everything; the C2 layer shapes (ffjord_tabular.py:229-232: 43 -> 860 -> 860 -> 43, B = 1000)
SHAPES = [(128, 128, 32), (128, 64, 256), (200, 150, 100), (1000, 860, 860), (1000, 860, 43), (1000, 43, 860), (860, 860, 3072)]


@pytest.mark.parametrize("bn", [128, 64])
@pytest.mark.parametrize("shape", SHAPES)
def test_tf32x3_gemm_matches_fp64(cuda, shape, bn):
    import easy_neural_ode_b200 as eno
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    ka = (K + 3) // 4 * 4
    A = torch.randn((M, ka), generator=g)[:, :K].to(cuda)       # row stride ka (a multiple of 4 elements, TMA needs 16 bytes)
    B = torch.randn((N, ka), generator=g)[:, :K].to(cuda)
    want = A.double() @ B.double().T
    got = _gemm(eno, A, B, bn)
    scale = (A.double().abs() @ B.double().abs().T)                # per-entry condition: sum_k |a||b|
    err = ((got.double() - want).abs() / scale).max().item()
    ref32 = ((A @ B.T).double() - want).abs().div(scale).max().item()   # what cuBLAS fp32 does on the same data
    assert torch.isfinite(got).all()
    print(f"tf32x3 {shape} bn={bn}: max err / sum|a||b| = {err:.3e} (fp32 cuBLAS on the same data: {ref32:.3e})")
    assert err < 1e-6, (err, ref32)


def test_split_k_is_deterministic_and_exact(cuda):
    import easy_neural_ode_b200 as eno
    g = torch.Generator().manual_seed(5)
    A = torch.randn((300, 3072), generator=g).to(cuda)
    B = torch.randn((172, 3072), generator=g).to(cuda)
    want = A.double() @ B.double().T
    a = _gemm(eno, A, B, 128, splits=3)
    b = _gemm(eno, A, B, 128, splits=3)
    assert torch.equal(a, b)
    assert ((a.double() - want).abs() / (A.double().abs() @ B.double().abs().T)).max().item() < 1e-6


def test_single_pass_tf32_would_not_be_enough(cuda):
    """Documents why three passes: the hi parts alone (what one tf32 MMA sees) are ~1e-3 off."""
    g = torch.Generator().manual_seed(1)
    A = torch.randn((128, 512), generator=g).to(cuda)
    B = torch.randn((128, 512), generator=g).to(cuda)
    hi = lambda x: (x.view(torch.int32) & -8192).view(torch.float32)
    want = A.double() @ B.double().T
    one = hi(A).double() @ hi(B).double().T
    scale = A.double().abs() @ B.double().abs().T
    assert ((one - want).abs() / scale).max().item() > 1e-5


@pytest.mark.parametrize("bn", [64, 128])
@pytest.mark.parametrize("shape", [(128, 64, 32), (64, 64, 256), (200, 150, 100), (784, 100, 300), (64, 64, 4096)])
def test_tf32x3_gemm_mn_major_operands_match_fp64(cuda, shape, bn):
    """MN-major operand descriptors (the contraction index is the slow one of both operands): C = A^T B with A [K, M], B [K, N]
    as they lie in memory - no transposed copies.  Same bar as the K-major mode."""
    import easy_neural_ode_b200 as eno
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 5 + N * 11 + K)
    ma, na = (M + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.randn((K, ma), generator=g).to(cuda)[:, :M]       # row stride ma / na (multiples of 4 elements: TMA needs 16 bytes)
    B = torch.randn((K, na), generator=g).to(cuda)[:, :N]
    want = A.double().T @ B.double()
    L = eno._cabi.lib()
    h = eno._cabi.ctx(cuda.index)
    for splits in (1, 3):
        out = torch.full((M, N), float("nan"), device=cuda)
        st = C.c_void_p(torch.cuda.current_stream(cuda).cuda_stream)
        rc = L.eno_gemm_tf32x3_mn(h, M, N, K, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0),
                                  C.c_void_p(out.data_ptr()), out.stride(0), bn, splits, st)
        eno._cabi.check(rc, h, "eno_gemm_tf32x3_mn")
        scale = A.double().abs().T @ B.double().abs()
        err = ((out.double() - want).abs() / scale).max().item()
        print(f"tf32x3 MN-major {shape} bn={bn} splits={splits}: max err / sum|a||b| = {err:.3e}")
        assert torch.isfinite(out).all()
        assert err < 1e-6, err
