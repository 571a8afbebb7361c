#!/usr/bin/env python3
"""Per-phase timing of the data-parallel step on N GPUs (torchrun, one rank per GPU). For each mode it runs un-graphed
steps with the library's per-launch CUDA-event marks (rank 0 printed; every rank executes the same sequence) and then
times graph replays of the whole step. Modes: local backward only, peer step without / with the copy-engine overlap."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
from ultrafnd_git_b200._lib import check

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 128
torch.manual_seed(0)
f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
f.train(); c.train()
step = FusedStep(f, c, B, use_graph=False, dp_group=dist.group.WORLD)
eng, plan, lib = step.engine, step.plan, step.engine.lib
lib.fnd_set_loss_scale(plan.handle, 1.0 / (B * world), eng.stream_ptr())
step.load_batch({k: v.to(dev) for k, v in bench.synth_batch(B, 1 + rank).items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ctypes.create_string_buffer(64 * 64)
ms = (ctypes.c_float * 64)()
cnt = ctypes.c_int()


def marks(fn, reps=20):
    acc = {}
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize(); dist.barrier()
        check(lib.fnd_profile_begin(plan.handle, eng.stream_ptr()), "begin")
        fn()
        check(lib.fnd_profile_end(plan.handle, eng.stream_ptr(), names, ms, 64, ctypes.byref(cnt)), "end")
        for j in range(cnt.value):
            nm = names.raw[64 * j:64 * j + 64].split(b"\0")[0].decode()
            acc[nm] = acc.get(nm, 0.0) + ms[j] * 1e3 / reps
    return acc


def timed(fn, reps=50):
    ts = []
    for _ in range(reps):
        flush.zero_()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for ov in (False, "defer"):
    step.dp_overlap = False
    step.dp_defer = ov == "defer"
    step.use_graph = False
    for _ in range(3):
        step.train_step_dp()
    a = marks(step.train_step_dp)
    if rank == 0:
        tail = {k: round(v, 1) for k, v in a.items() if k.startswith(("dp_", "wgrad", "grad_norm", "finalize"))}
        print(f"[world {world}] overlap={ov}: un-graphed marks (us, rank 0): sum {sum(a.values()):.1f}; tail {tail}")
    step.use_graph = True
    step._graphs.clear()
    for _ in range(3):
        step.train_step_dp()
    t = timed(step.train_step_dp)
    if rank == 0:
        print(f"[world {world}] overlap={ov}: graph replay {t:.1f} us/step (max over ranks, median of 50)")
step.use_graph = True
step._graphs.clear()
for _ in range(3):
    step.train_fwd_bwd()
t = timed(step.train_fwd_bwd)
if rank == 0:
    print(f"[world {world}] local forward+backward only (no optimizer): {t:.1f} us")
plan.check_error()
dist.destroy_process_group()
