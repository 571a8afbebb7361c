#!/usr/bin/env python3
"""Marginal in-graph cost of every kernel of the training step: time CUDA-graph replays of the first n launches of
fnd_train_step for n = 1..18 (L2 flushed between replays) and print the differences. Unlike per-kernel events this
sees the step exactly as the bench does (PDL overlap, graph launch) — the number next to a kernel is what removing it
(and nothing else) would save."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
import bench

NAMES = ["prep", "gemm_proj", "gemm_qkv", "assemble_fwd", "gemm_fuse0", "gemm_fuse1", "gemm_pre0", "gemm_pre1", "head",
         "dgrad_pre1", "dgrad_pre0", "dgrad_fuse1", "dgrad_fuse0", "assemble_bwd", "dgrad_qkv", "wgrad_all+fin", "adamw"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = 60
torch.manual_seed(0)
f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
f.train(); c.train()
step = FusedStep(f, c, B, use_graph=True)
step.load_batch({k: v.cuda() for k, v in bench.synth_batch(B, 1).items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib, h = step.engine.lib, step.plan.handle
prev = 0.0
print(f"batch {B}; cumulative / marginal microseconds per graph replay (median of {reps}, L2 flushed)")
for n in range(1, len(NAMES) + 1):
    lib.fnd_debug_set_launch_limit(h, n)
    step._graphs.clear()
    for _ in range(3):
        step.train_step()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step.train_step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{n:2d} {NAMES[n - 1]:13s} cum {med:8.1f}  marginal {med - prev:7.1f}")
    prev = med
lib.fnd_debug_set_launch_limit(h, -1)
