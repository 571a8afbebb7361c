#!/usr/bin/env python3
"""Localise a data-parallel reduction error: compare every rank's reduced slice (gred) with an NCCL all-reduce of the
same local gradients, segment by segment (torchrun, one rank per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = 32
torch.manual_seed(3)
f, c = CrossModalTransformer(precision=precision), DeepTruthClassifier(precision=precision)
for m in list(f.modules()) + list(c.modules()):
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
step = FusedStep(f, c, B, precision=precision, use_graph=False, dp_group=dist.group.WORLD)
eng, plan = step.engine, step.plan
eng.lib.fnd_set_loss_scale(plan.handle, 1.0 / (B * world), eng.stream_ptr())
step.load_batch({k: v.to(dev) for k, v in bench.synth_batch(B, 10 + rank).items()})
init = eng.params.clone(); sh0 = eng.shadow_hi.clone(); sl0 = eng.shadow_lo.clone() if eng.shadow_lo is not None else None
step.train_fwd_bwd()
torch.cuda.synchronize()
base = eng.grads.clone()
dist.all_reduce(base)
print(f"[rank {rank}] plain train_fwd_bwd: norm of all-reduced grads {float(base.double().norm()):.8f}", flush=True)
for overlap in (False, True, "graph"):
    eng.params.copy_(init); eng.shadow_hi.copy_(sh0); eng.adam_m.zero_(); eng.adam_v.zero_()
    if sl0 is not None:
        eng.shadow_lo.copy_(sl0)
    torch.cuda.synchronize(); dist.barrier()
    if overlap == "graph":
        step.dp_overlap = True; step.use_graph = True; step._graphs.clear()
    else:
        step.dp_overlap = overlap
    step.train_step_dp()
    torch.cuda.synchronize()
    ref = eng.grads.clone()
    dist.all_reduce(ref)
    print(f"[rank {rank} overlap={overlap}] all-reduced grads vs plain: max diff {float((ref - base).abs().max()):.3e}; "
          f"norm {float(ref.double().norm()):.8f}; state grad_norm {plan.state()['grad_norm']:.8f}", flush=True)
    gred = eng.symm["gred"]
    off = 0
    for si, (lo, hi) in enumerate(eng.shard_ranges(rank)):
        mine = gred[off:off + hi - lo]
        want = ref[lo:hi]
        err = (mine - want).abs().max().item()
        scale = want.abs().max().item()
        bad = ((mine - want).abs() > 1e-3 * scale + 1e-12).nonzero().flatten()
        print(f"[rank {rank} overlap={overlap}] seg {si} [{lo},{hi}) max err {err:.3e} (scale {scale:.3e}) bad {bad.numel()}"
              + (f" first bad idx {int(bad[0]) + lo} last {int(bad[-1]) + lo}" if bad.numel() else ""), flush=True)
        off += hi - lo
    dist.barrier()
plan.check_error()
dist.destroy_process_group()
