#!/usr/bin/env python3
"""Event-timed back-to-back launches of one GEMM shape (hot caches) vs launches separated by an L2 flush."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultrafnd_git_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
def run(M, N, K, a_mn, b_mn, bn, splits):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    C = torch.empty(M, N, device=dev)
    nb = lib.fnd_gemm_scratch_bytes(M, N, bn, splits)
    scratch = torch.zeros(nb + 256, dtype=torch.uint8, device=dev)
    sp = (scratch.data_ptr() + 255) // 256 * 256
    st = torch.cuda.current_stream().cuda_stream
    def call(reps):
        lib.fnd_gemm_bf16_probe(A.data_ptr(), A.data_ptr(), A.shape[1], a_mn, B.data_ptr(), B.data_ptr(), B.shape[1], b_mn,
                                C.data_ptr(), N, M, N, K, bn, splits, 1, sp, nb, st, None, reps)
    call(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); call(200); e1.record(); torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) * 1e3 / 200
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for i in range(10):
        flush.zero_(); torch.cuda.synchronize()
        e0.record(); call(1); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) * 1e3
    print(f"M{M} N{N} K{K} a{a_mn} b{b_mn} bn{bn} s{splits}: back-to-back {hot:.2f} us/launch; after L2 flush (incl. memset+sync of the call) {tot/10:.2f} us")
for c in [(128, 512, 512, 0, 0, 64, 1), (128, 512, 512, 0, 1, 64, 1), (128, 1024, 8192, 0, 0, 64, 16), (128, 1024, 8192, 0, 0, 64, 8), (128, 1024, 8192, 0, 0, 128, 16), (512, 8192, 128, 1, 1, 128, 1), (1024, 8192, 128, 1, 1, 128, 1)]:
    run(*c)
