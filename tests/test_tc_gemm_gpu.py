"""GPU: the tcgen05 3xTF32 GEMM vs a float64 matmul (fp32-level accuracy is the contract)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(fpv, x):
    L = fpv._lib.lib()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    fpv._lib.check(L.fpv_split_tf32(fpv._lib.ptr(x), x.numel(), fpv._lib.ptr(hi), fpv._lib.ptr(lo), fpv._lib.stream_ptr()))
    return hi, lo


def _gemm(fpv, A, B, M, N, K, ksplit=1):
    """A [M, lda], B [N, ldb] (padded pitches), logical K columns."""
    L = fpv._lib.lib()
    ah, al = _split(fpv, A)
    bh, bl = _split(fpv, B)
    assert torch.equal(ah + al, A) or (ah + al - A).abs().max() <= 2e-7 * A.abs().max()
    C = torch.full((M, N + 3), float("nan"), device=A.device)
    ws = fpv._lib.workspace(L.fpv_tc_gemm_workspace_bytes(M, N, ksplit), A.device)
    fpv._lib.check(L.fpv_tc_gemm_3xtf32(fpv._lib.ptr(ah), fpv._lib.ptr(al), A.stride(0), fpv._lib.ptr(bh), fpv._lib.ptr(bl),
                                        B.stride(0), M, N, K, fpv._lib.ptr(C), C.stride(0), ksplit, fpv._lib.ptr(ws),
                                        ws.numel(), fpv._lib.stream_ptr()), "fpv_tc_gemm_3xtf32")
    torch.cuda.synchronize()
    assert torch.isnan(C[:, N:]).all()          # nothing written outside the logical columns
    return C[:, :N]


@pytest.mark.parametrize("M,N,K,ksplit", [(128, 128, 32, 1), (128, 128, 512, 1), (300, 1000, 512, 1),
                                          (300, 31425, 512, 1), (300, 512, 31425, 12), (77, 130, 100, 3)])
def test_tc_gemm_matches_float64(fpv, cuda_dev, M, N, K, ksplit):
    g = torch.Generator().manual_seed(M + N + K)
    Kp = (K + 3) // 4 * 4
    A = torch.zeros(M, Kp)
    B = torch.zeros(N, Kp)
    A[:, :K] = torch.randn(M, K, generator=g)
    B[:, :K] = torch.randn(N, K, generator=g) * 0.01
    A[:, K:] = 7.0                                # pitch padding must be ignored (TMA bounds, not values)
    B[:, K:] = 7.0
    ref = A[:, :K].double() @ B[:, :K].double().t()
    C = _gemm(fpv, A.to(cuda_dev), B.to(cuda_dev), M, N, K, ksplit).cpu().double()
    scale = (A[:, :K].abs().double() @ B[:, :K].abs().double().t())
    err = ((C - ref).abs() / scale).max().item()
    assert err < 2e-6, err
    C2 = _gemm(fpv, A.to(cuda_dev), B.to(cuda_dev), M, N, K, ksplit).cpu().double()
    assert torch.equal(C, C2)                      # deterministic, including split-K
