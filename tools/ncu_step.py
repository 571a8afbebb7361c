#!/usr/bin/env python3
"""Plain (un-graphed) training steps at the bench configuration, for Nsight Compute:
3 warm-up steps then 2 measured steps of fnd_train_step at batch 128, bf16 mode, L2 flushed before each step.
17 library kernels per step; with `-k regex:fnd_|prep_|assemble_|head_|adamw_ -s 51 -c 17` ncu sees exactly one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
f.train(); c.train()
step = FusedStep(f, c, B, use_graph=False)
step.load_batch({k: v.cuda() for k, v in bench.synth_batch(B, 1).items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(5):
    flush.zero_()
    torch.cuda.synchronize()
    step.train_step()
torch.cuda.synchronize()
step.plan.check_error()
print("ok", step.plan.state()["loss"])
