#!/usr/bin/env python3
"""Phase stamps of the fused head kernel inside a real training step (FND_DEBUG_STAMPS=1)."""
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
t = step.plan.buffer("dbg", torch.int64, (4096, 8))[:16].cpu().double()
d = (t - t[:, :1]) / 1.965e3
names = ["start", "staged", "setup", "fwd_dots", "fwd_done", "bwd_trees", "bwd_done", "end"]
print("head kernel, median us since CTA start:", {n: round(float(d[:, i].median()), 2) for i, n in enumerate(names)})
print("max end", float(d[:, 7].max()))
