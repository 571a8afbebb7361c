"""GPU parity of the tcgen05 grouped GEMM (C-ABI fnd_gemm_bf16) against torch fp32 matmul."""
import ctypes

import pytest
import torch

from ultrafnd_git_b200 import _lib

pytestmark = pytest.mark.gpu


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def run_gemm(A, B, a_mn, b_mn, bn, splits, ncombo):
    """A: [M,K] fp32, B: [N,K] fp32 logical. Stores operands in the requested memory order."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    a_mem = A.t().contiguous() if a_mn else A.contiguous()
    b_mem = B.t().contiguous() if b_mn else B.contiguous()
    a_hi, a_lo = _split(a_mem)
    b_hi, b_lo = _split(b_mem)
    C = torch.full((M, N), float("nan"), device=A.device, dtype=torch.float32)
    nbytes = lib.fnd_gemm_scratch_bytes(M, N, bn, splits)
    scratch = torch.zeros(nbytes + 256, dtype=torch.uint8, device=A.device)
    sp = (scratch.data_ptr() + 255) // 256 * 256
    st = lib.fnd_gemm_bf16(a_hi.data_ptr(), a_lo.data_ptr(), a_mem.shape[1], int(a_mn),
                           b_hi.data_ptr(), b_lo.data_ptr(), b_mem.shape[1], int(b_mn),
                           C.data_ptr(), N, M, N, K, bn, splits, ncombo,
                           sp, nbytes, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if ncombo == 1:
        ref = a_hi.float().t() @ b_hi.float() if False else None
        Ar = (a_hi.float().t() if a_mn else a_hi.float())
        Br = (b_hi.float().t() if b_mn else b_hi.float())
        ref = Ar.double() @ Br.double().t()
    else:
        ref = A.double() @ B.double().t()
    return st, C, ref.float()


CASES = [
    # (M, N, K, a_mn, b_mn, bn, splits)
    (128, 128, 64, 0, 0, 128, 1),
    (128, 128, 256, 0, 0, 128, 1),
    (128, 512, 768, 0, 0, 64, 1),
    (128, 1024, 8192, 0, 0, 64, 8),
    (4, 512, 128, 0, 0, 64, 1),
    (100, 512, 512, 0, 0, 32, 2),
    (300, 1024, 1024, 0, 0, 128, 3),
    (128, 128, 64, 0, 1, 128, 1),
    (128, 512, 1024, 0, 1, 64, 1),
    (260, 8192, 1024, 0, 1, 128, 1),
    (128, 128, 64, 1, 1, 128, 1),
    (512, 768, 128, 1, 1, 128, 1),
    (1024, 8192, 128, 1, 1, 128, 1),
    (512, 512, 4, 1, 1, 64, 1),
    (24, 512, 300, 1, 1, 128, 1),
    (128, 128, 128, 1, 0, 128, 1),
    # narrow tiles / deep rings / distributed split-K fix-up (the configurations the batch-128 step uses)
    (128, 512, 512, 0, 0, 16, 1),
    (128, 2560, 768, 0, 0, 32, 1),
    (128, 1024, 8192, 0, 0, 128, 16),
    (128, 512, 1024, 0, 0, 16, 4),
    (128, 64, 2048, 0, 0, 16, 16),
    (77, 512, 1536, 0, 1, 64, 3),
    (128, 512, 4096, 0, 0, 64, 7),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "M%d_N%d_K%d_a%d_b%d_bn%d_s%d" % c)
@pytest.mark.parametrize("ncombo", [1, 3])
def test_gemm_parity(case, ncombo):
    M, N, K, a_mn, b_mn, bn, splits = case
    g = torch.Generator(device="cpu").manual_seed(1234 + M + N + K)
    # MN-major operands need 16-byte-aligned rows: pad the row dimension in memory when required
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    if a_mn and M % 8:
        pytest.skip("MN-major A needs M % 8 == 0 in this direct test")
    if b_mn and N % 8:
        pytest.skip("MN-major B needs N % 8 == 0 in this direct test")
    if (not a_mn or not b_mn) and K % 8 and not (a_mn and b_mn):
        pytest.skip("K-major operands need K % 8 == 0")
    st, C, ref = run_gemm(A, B, a_mn, b_mn, bn, splits, ncombo)
    assert st == 0, f"status {st}"
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 1e-4
    print(f"\n[gemm] {case} ncombo={ncombo} max_abs_err={err:.3e} ref_max={scale:.3e} rel={err / scale:.3e}")
    assert torch.isfinite(C).all()
    assert err / scale < tol
