#!/usr/bin/env python3
"""Data-parallel ForensicTrainer on N GPUs (torchrun): fit on a synthetic cache, then check that every rank holds the
same model (bf16 shadows bit-identical; fp32 master identical after the gather) and that it learned."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from ultrafnd_git_b200.trainer import ForensicTrainer, TrainConfig, synthetic_cache


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out_dir = os.path.join(tempfile.gettempdir(), "fnd_dp_trainer_check")
    # n and batch size deliberately NOT divisible by the world size; 449 training rows = 7 x 64 + 1, so in the last global
    # batch every rank but 0 has an empty shard and runs the zero-weight padding step
    cache = synthetic_cache(n=642, seed=3)
    assert len(cache["split"][0]) == 449
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir=out_dir, batch_size=64, epochs=3, lr=5e-4)
    tr = ForensicTrainer(cfg, cache=cache)
    assert tr.dp_peer, "expected the peer-memory optimizer step with the NCCL backend"
    best = tr.fit()
    eng = tr.engine
    eng.gather_master()
    dev = eng.device

    def same_everywhere(t):
        x = t.contiguous().view(torch.int16).to(torch.int64) if t.dtype == torch.bfloat16 else t.contiguous().view(torch.int32).to(torch.int64)
        chk = torch.stack([x.sum(), (x * (torch.arange(x.numel(), device=dev) % 8191)).sum()])
        parts = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(parts, chk)
        return all(torch.equal(p, parts[0]) for p in parts)
    ok = same_everywhere(eng.shadow_hi) and same_everywhere(eng.params[:eng.n_hot])
    res = tr.test()
    # every rank must report the same numbers (metrics are gathered, never rank-local)
    t = torch.tensor([best, res["test_auc"], res["test_loss"]], dtype=torch.float64, device=dev)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    ok = ok and all(torch.equal(p, parts[0]) for p in parts)
    if rank == 0:
        print(f"[dp_trainer_check world={world}] best val auc {best:.3f}, test auc {res['test_auc']:.3f}, replicas identical: {ok}")
    good = ok and best > 0.9 and res["test_auc"] > 0.85
    flag = torch.tensor([1 if good else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DP TRAINER CHECK", "OK" if int(flag.item()) else "FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
