#!/usr/bin/env python3
"""In-kernel phase stamps of every GEMM/head launch inside a real (un-graphed, L2-flushed) training step."""
import os, sys
os.environ["FND_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
import bench
torch.manual_seed(0)
B = 128
f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
f.train(); c.train()
step = FusedStep(f, c, B, use_graph=False)
step.load_batch({k: v.cuda() for k, v in bench.synth_batch(B, 1).items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    flush.zero_()
    step.train_step()
torch.cuda.synchronize()
names = ["gemm_proj", "gemm_qkv", "gemm_fuse0", "gemm_fuse1", "gemm_pre0", "gemm_pre1", "head", "dgrad_pre1", "dgrad_pre0",
         "dgrad_fuse1", "dgrad_fuse0", "dgrad_qkv", "wgrad_all"]
buf = step.plan.buffer("dbg", torch.int64, (40, 1024, 8)).cpu().double()
lab = ["setup", "first_full", "mma_issued", "accum_rdy", "splitk", "epi_done"]
for i, n in enumerate(names):
    g = int((buf[i, :, 0] != 0).sum())          # CTAs that stamped (grid <= 1024 is recorded)
    t = buf[i, :g]
    d = (t - t[:, :1]) / 1.965e3
    if n == "head":
        print(f"{n:12s} grid {g:4d} median us:", [round(float(d[:, j].median()), 2) for j in range(1, 8)])
    else:
        print(f"{n:12s} grid {g:4d} median us:", {l: round(float(d[:, j + 1].median()), 2) for j, l in enumerate(lab)},
              "max epi", round(float(d[:, 6].max()), 2))

# ---- wgrad launch: per-problem CTA durations (tile CTAs only) ----
i = names.index("wgrad_all")
g = int((buf[i, :, 0] != 0).sum())
t = buf[i, :g]
dur = (t[:, 7] - t[:, 0]) / 1.965e3
seg = [("pre.3", 16), ("pre.0", 16), ("dA", 4), ("fuse1", 32), ("fuse0", 512), ("qkv", 144), ("proj", 56)]
o = 0
print("wgrad per-problem CTA duration us (median / p90 / max) and phase medians [setup, first_full, mma, accum, epi0, epi_done]")
for nm, n in seg:
    d = dur[o:o + n]
    ph = ((t[o:o + n, 1:7] - t[o:o + n, :1]) / 1.965e3).median(0).values
    print(f"  {nm:6s} n={n:4d} {float(d.median()):6.2f} {float(d.quantile(0.9)):6.2f} {float(d.max()):6.2f}  ", [round(float(x), 2) for x in ph])
    o += n
