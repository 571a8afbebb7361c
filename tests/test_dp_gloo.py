"""World-size-2 (gloo, CPU) test of the data-parallel host logic: batch sharding + loss scaling + SUM all-reduce of
the gradients reproduce the single-process gradient, and the broadcast makes replicas identical."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fnd_oracle as O
    from ultrafnd_git_b200.trainer import shard_indices
    torch.set_num_threads(2)
    fus, clf = O.init_params(42 + rank)          # deliberately different replicas before the broadcast
    for d in (fus, clf):
        for k in sorted(d):
            dist.broadcast(d[k], src=0)
    O.perturb_node_head(clf)
    B = 10                                        # ragged: 10 samples over 2 ranks, an odd global batch below
    batch = O.make_batch(B + 1, seed=5)
    gidx = torch.arange(B + 1)
    local = shard_indices(gidx, rank, world)
    lb = {k: v[local] for k, v in batch.items()}
    fk, ck = O.trainable_keys()
    fl = {k: v.clone().requires_grad_(k in fk) for k, v in fus.items()}
    cl = {k: v.clone().requires_grad_(k in ck) for k, v in clf.items()}
    out = O.model_forward(fl, cl, lb, dropout=0.0)
    rows = torch.nn.functional.cross_entropy(out["logits"], lb["label"], reduction="none")
    (rows.sum() * (1.0 / (B + 1))).backward()     # per-rank loss scale = 1 / len(global batch)
    flat = torch.cat([fl[k].grad.flatten() for k in fk] + [cl[k].grad.flatten() for k in ck])
    dist.all_reduce(flat)                         # SUM
    covered = torch.zeros(B + 1)
    covered[local] = 1
    dist.all_reduce(covered)
    if rank == 0:
        _, gf, gc = O.loss_and_grads(fus, clf, batch, dropout=0.0)
        ref = torch.cat([gf[k].flatten() for k in fk] + [gc[k].flatten() for k in ck])
        ret["err"] = O.rel_err(flat, ref)
        ret["covered"] = bool((covered == 1).all())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_single_process_gradients():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["covered"], "shards must be disjoint and exhaustive"
    assert ret["err"] < 1e-5, ret["err"]
