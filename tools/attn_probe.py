#!/usr/bin/env python3
"""Timing of the fused attention forward / backward kernels alone at the stress shape (CUDA events, warm-up, median).
usage: python tools/attn_probe.py [B] [Lq] [Lk] [heads] [reps]      (the ncu captures in profiles/ wrap this command)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ultrafnd_git_b200 import seq_ops as S


def med(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    Lk = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    d = H * 64
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv_q = torch.randn(B * Lq, 3 * d, device=dev, generator=g).bfloat16()
    qkv_k = torch.randn(B * Lk, 3 * d, device=dev, generator=g).bfloat16()
    d_o = torch.randn(B * Lq, d, device=dev, generator=g).bfloat16()
    o = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Lq, device=dev, dtype=torch.float32)
    dqkv_q = torch.zeros_like(qkv_q)
    dqkv_k = torch.zeros_like(qkv_k)
    err = S.new_err_flag(dev)
    flops = 4.0 * B * Lq * Lk * d
    fwd = lambda: S.coattn_forward(qkv_q, qkv_k, qkv_k, B, H, Lq, Lk, 0, d, 2 * d, out=o, lse=lse, err=err)
    bwd = lambda: S.coattn_backward(qkv_q, qkv_k, qkv_k, o, d_o, lse, B, H, Lq, Lk, dqkv_q, dqkv_k, dqkv_k,
                                    q_col0=0, k_col0=d, v_col0=2 * d, dq_col0=0, dk_col0=d, dv_col0=2 * d, err=err)
    ms = med(fwd, reps)
    print(f"attn fwd  B={B} H={H} Lq={Lq} Lk={Lk}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s")
    if os.environ.get("ATTN_PROBE_SWEEP"):
        for pp in (0, 1):
            for poly in (0, 1, 2, 3):
                os.environ["FND_ATTN_POLY"], os.environ["FND_ATTN_PINGPONG"] = str(poly), str(pp)
                ms = med(fwd, reps)
                print(f"   pingpong={pp} poly={poly}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s")
        del os.environ["FND_ATTN_POLY"], os.environ["FND_ATTN_PINGPONG"]
    if not os.environ.get("ATTN_PROBE_FWD_ONLY"):
        ms = med(bwd, reps)
        # 7 GEMMs of 2 * Lq * Lk * 64 per head: the algorithmic backward is 5 (2.5x the forward)
        print(f"attn bwd  B={B} H={H} Lq={Lq} Lk={Lk}: {ms * 1e3:8.1f} us  {2.5 * flops / ms / 1e9:8.1f} TFLOP/s (algorithmic 2.5x fwd FLOPs)")
    torch.cuda.synchronize()
    print("err", int(err.item()))


if __name__ == "__main__":
    main()
