#!/usr/bin/env python3
"""Kernel-level timing of the sequence front-end building blocks at the stress shape (CUDA events, warm-up, median).
usage: python tools/seq_probe.py [B] [Lt] [Lf] [d]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ultrafnd_git_b200 import seq_ops as S


def timeit(fn, reps=int(os.environ.get('SEQ_PROBE_REPS', '20')), warm=int(os.environ.get('SEQ_PROBE_WARM', '3'))):
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
    Lt = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    Lf = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    d = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    H = d // 64
    dev = torch.device("cuda")
    err = S.new_err_flag(dev)
    g = torch.Generator(device="cuda").manual_seed(0)
    xt = torch.randn(B * Lt, d, device=dev, generator=g).bfloat16()
    xf = torch.randn(B * Lf, d, device=dev, generator=g).bfloat16()
    w3 = (torch.randn(3 * d, d, device=dev, generator=g) / math.sqrt(d)).bfloat16()
    wo = (torch.randn(d, d, device=dev, generator=g) / math.sqrt(d)).bfloat16()
    b3, bo = torch.zeros(3 * d, device=dev), torch.zeros(d, device=dev)
    qkv_t = torch.empty(B * Lt, 3 * d, device=dev, dtype=torch.bfloat16)
    qkv_f = torch.empty(B * Lf, 3 * d, device=dev, dtype=torch.bfloat16)
    at = torch.empty(B * Lt, d, device=dev, dtype=torch.bfloat16)
    af = torch.empty(B * Lf, d, device=dev, dtype=torch.bfloat16)
    yt = torch.empty(B * Lt, d, device=dev, dtype=torch.bfloat16)
    ones = torch.ones(d, device=dev); zeros = torch.zeros(d, device=dev)
    rows = []

    def rec(name, ms, flops=0.0, nbytes=0.0):
        rows.append((name, ms, flops / ms / 1e9 if flops else 0.0, nbytes / ms / 1e6 if nbytes else 0.0))

    ms = timeit(lambda: S.linear(xt, w3, b3, out=qkv_t, err=err)); rec("in_proj text  [%d x %d x %d]" % (B * Lt, 3 * d, d), ms, 2.0 * B * Lt * 3 * d * d)
    ms = timeit(lambda: S.linear(xf, w3, b3, out=qkv_f, err=err)); rec("in_proj frames[%d x %d x %d]" % (B * Lf, 3 * d, d), ms, 2.0 * B * Lf * 3 * d * d)
    # comparator only (never on the product path): the library GEMM on the same shapes, bias / residual not included
    ms = timeit(lambda: torch.matmul(xt, w3.t(), out=qkv_t)); rec("  [cuBLAS via torch.matmul, same shape, no bias]", ms, 2.0 * B * Lt * 3 * d * d)
    ms = timeit(lambda: torch.matmul(xf, w3.t(), out=qkv_f)); rec("  [cuBLAS frames shape]", ms, 2.0 * B * Lf * 3 * d * d)
    ms = timeit(lambda: torch.matmul(xt, wo.t(), out=yt)); rec("  [cuBLAS out_proj shape, no residual]", ms, 2.0 * B * Lt * d * d)
    ms = timeit(lambda: S.coattn_forward(qkv_t, qkv_f, qkv_f, B, H, Lt, Lf, 0, d, 2 * d, out=at, err=err)); rec("attn text<-frames", ms, 4.0 * B * Lt * Lf * d)
    ms = timeit(lambda: S.coattn_forward(qkv_f, qkv_t, qkv_t, B, H, Lf, Lt, 0, d, 2 * d, out=af, err=err)); rec("attn frames<-text", ms, 4.0 * B * Lt * Lf * d)
    ms = timeit(lambda: S.linear(at, wo, bo, resid=xt, out=yt, err=err)); rec("out_proj text +resid", ms, 2.0 * B * Lt * d * d)
    ms = timeit(lambda: S.linear(at, wo, bo, out=yt, err=err)); rec("out_proj text (no resid)", ms, 2.0 * B * Lt * d * d)
    ms = timeit(lambda: S.layernorm(yt, ones, zeros, out=at)); rec("layernorm text", ms, 0.0, 4.0 * B * Lt * d)
    ms = timeit(lambda: S.layernorm(yt, ones, zeros, out=at, resid=xt)); rec("layernorm text (x + resid)", ms, 0.0, 6.0 * B * Lt * d)
    ms = timeit(lambda: S.masked_mean_pool(at, B, Lt)); rec("pool text", ms, 0.0, 2.0 * B * Lt * d)
    xin = torch.randn(B * Lt, 768, device=dev, generator=g)
    ms = timeit(lambda: S.cast_bf16(xin)); rec("cast text fp32->bf16", ms, 0.0, 6.0 * B * Lt * 768)
    if os.environ.get("SEQ_PROBE_STAMPS"):
        from ultrafnd_git_b200 import _lib
        lib = _lib.load()
        st = torch.zeros(2 * 148 * 8 + 64, dtype=torch.int64, device=dev)
        lib.fnd_seq_debug_attn_stamps(st.data_ptr())
        S.coattn_forward(qkv_t, qkv_f, qkv_f, B, H, Lt, Lf, 0, d, 2 * d, out=at, err=err)
        torch.cuda.synchronize()
        lib.fnd_seq_debug_attn_stamps(None)
        v = st[:2 * 148 * 8].view(-1, 8).double().cpu()
        nb = B * H * ((Lt + 127) // 128) * ((Lf + 63) // 64) / (2 * 148)
        names = ["wait S", "tmem ld S", "mask+max+xchg", "exp2+pack", "P store+fence", "prev PV add", "rescale", "loop/epilogue"]
        print(f"attention phase cycles per key block (softmax warp 2 lane 0, mean over {v.shape[0]} CTAs, {nb:.1f} blocks/CTA):")
        for i, n in enumerate(names):
            print(f"   {n:16s} {float(v[:, i].mean()) / nb:8.1f}")
        print(f"   {'total':16s} {float(v.sum(1).mean()) / nb:8.1f}")
    torch.cuda.synchronize()
    print(f"seq_probe B={B} Lt={Lt} Lf={Lf} d={d} heads={H}  err={int(err.item())}")
    for name, ms, tf, gbs in rows:
        print(f"  {name:44s} {ms * 1e3:9.1f} us  {tf:8.1f} TFLOP/s  {gbs:8.1f} GB/s")
    tot = sum(r[1] for r in rows[:5]) + rows[4][1] * Lf / Lt + rows[6][1] * (1 + Lf / Lt)
    fl = B * 2.0 * (2.0 * (2.0 * Lt * d * d + 2.0 * Lf * d * d + 2.0 * Lt * Lf * d))
    print(f"  one bidirectional layer (sum of parts): {tot * 1e3:.1f} us, {fl / tot / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
