#!/usr/bin/env python3
"""clock64 totals of the CTA-pair projection GEMM (csrc/fnd_seq_gemm2.cuh) through fnd_seq_debug_gemm_stamps: where the
MMA-issuing warp, the TMA producer and the epilogue warps spend their cycles.  python tools/gemm2_stamps.py [M N K resid]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ultrafnd_git_b200 import _lib, seq_ops as S


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    resid = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dev = torch.device("cuda")
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.zeros(N, device=dev)
    r = torch.randn(M, N, device=dev, generator=g).bfloat16() if resid else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    err = S.new_err_flag(dev)
    for _ in range(3):
        S.linear(a, w, b, resid=r, out=out, err=err)
    torch.cuda.synchronize()
    nc = lib.fnd_seq_pair_clusters()
    st = torch.zeros(16 * max(nc, 1), dtype=torch.int64, device=dev)
    lib.fnd_seq_debug_gemm_stamps(st.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S.linear(a, w, b, resid=r, out=out, err=err); e1.record()
    torch.cuda.synchronize()
    lib.fnd_seq_debug_gemm_stamps(None)
    v = st.view(-1, 16).double().cpu()
    it = float(v[:, 5].mean())
    print(f"pair GEMM [{M} x {N} x {K}] resid={resid}: {e0.elapsed_time(e1) * 1e3:.1f} us with stamps, {nc} clusters, {it:.0f} k-blocks per cluster, err={int(err.item())}")
    names = ["MMA loop total", "MMA wait accumulator drained", "MMA wait operands (full)", "MMA issue (fence, 4 UMMA, commits)",
             "producer wait free stage", "k-blocks", "epilogue loop total", "epilogue wait accumulator",
             "epilogue: store drained + barrier (+ residual landed)", "epilogue: tcgen05.ld x2 + wait", "epilogue: bias / residual / pack / st.shared",
             "epilogue: fence + barrier + TMA store issue", "MMA: fence + descriptor arithmetic", "MMA: __syncwarp after the elected block"]
    for i, n in enumerate(names):
        if i == 5:
            continue
        print(f"   {n:38s} {float(v[:, i].mean()):12.0f} cycles   {float(v[:, i].mean()) / it:8.1f} per k-block")


if __name__ == "__main__":
    main()
