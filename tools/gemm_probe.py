#!/usr/bin/env python3
"""Phase timing of the tcgen05 GEMM kernel: every CTA stamps clock64() at 8 points (fnd_gemm_bf16_probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultrafnd_git_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
CASES = [  # M, N, K, a_mn, b_mn, bn, splits
    (128, 64, 64, 0, 0, 64, 1), (128, 64, 512, 0, 0, 64, 1), (128, 64, 4096, 0, 0, 64, 1),
    (128, 512, 512, 0, 0, 64, 1), (128, 512, 512, 0, 1, 64, 1), (128, 1024, 8192, 0, 0, 64, 16),
    (512, 8192, 128, 1, 1, 128, 1),
]
for (M, N, K, a_mn, b_mn, bn, splits) in CASES:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    C = torch.empty(M, N, device=dev)
    nb = lib.fnd_gemm_scratch_bytes(M, N, bn, splits)
    scratch = torch.zeros(nb + 256, dtype=torch.uint8, device=dev)
    sp = (scratch.data_ptr() + 255) // 256 * 256
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    kb = (K + 63) // 64
    s_eff = min(splits, kb)
    kps = (kb + s_eff - 1) // s_eff
    grid = tiles * ((kb + kps - 1) // kps)
    stamps = torch.zeros(grid * 8, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for cold in (0, 1):
      for reps in (1, 1, 1):
        if cold:
            flush.zero_()
        lib.fnd_gemm_bf16_probe(A.data_ptr(), A.data_ptr(), A.shape[1], a_mn, B.data_ptr(), B.data_ptr(), B.shape[1], b_mn,
                                C.data_ptr(), N, M, N, K, bn, splits, 1, sp, nb, st, stamps.data_ptr(), 1)
      torch.cuda.synchronize()
      t = stamps.view(grid, 8).cpu().double()
      d = (t - t[:, :1]) / 1.965e3   # us at 1965 MHz
      names = ["start", "setup", "first_full", "mma_issued", "accum_ready", "splitk_done", "epi_done", "all_done"]
      print(f"case M{M} N{N} K{K} a{a_mn} b{b_mn} bn{bn} s{splits} grid {grid} {'COLD' if cold else 'hot '}: median us:",
            {n: round(float(d[:, i].median()), 2) for i, n in enumerate(names[1:7], 1)}, " max epi_done", round(float(d[:, 6].max()), 2))
