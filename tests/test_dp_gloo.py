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


def _zero1_worker(rank, world, port, ret):
    """Host-level restatement of csrc/fnd_dp.cuh with gloo: rank-order sum of the pieces a rank owns + AdamW on the
    slice + gather of the slices == AdamW on the all-reduced gradient (same element-wise arithmetic, so bitwise)."""
    import ctypes
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ultrafnd_git_b200 import _lib, engine as E
    lib = _lib.load()
    dims = E.Dims()
    cd = dims.to_c()
    h = ctypes.c_void_p()
    assert lib.fnd_plan_create(ctypes.byref(cd), 8, 0, ctypes.byref(h)) == 0
    n = lib.fnd_arena_hot_elems(ctypes.byref(cd))
    g = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(n, generator=g) * 1e-3                       # this rank's local gradient
    gp = torch.Generator().manual_seed(7)
    p0 = torch.randn(n, generator=gp) * 0.02                         # identical replicas
    m0, v0 = torch.zeros(n), torch.zeros(n)
    lr, b1, b2, eps, wd, max_norm, t = 2e-4, 0.9, 0.999, 1e-8, 1e-4, 5.0, 1

    def adamw(p, gr, m, v, coef):
        gr = gr * coef
        p = p * (1.0 - lr * wd)
        m = b1 * m + (1.0 - b1) * gr
        v = b2 * v + (1.0 - b2) * gr * gr
        denom = v.sqrt() / (1.0 - b2 ** t) ** 0.5 + eps
        return p - (lr / (1.0 - b1 ** t)) * (m / denom), m, v

    # ---- sharded path ----
    all_g = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(all_g, grads)                                   # stands in for the P2P pushes
    lo, hi = (ctypes.c_longlong * 3)(), (ctypes.c_longlong * 3)()
    assert lib.fnd_dp_shard_ranges(h, rank, world, lo, hi) == 3
    reduced = {}
    part = 0.0
    for s in range(3):
        acc = torch.zeros(hi[s] - lo[s])
        for r in range(world):                                       # rank order
            acc = acc + all_g[r][lo[s]:hi[s]]
        reduced[s] = acc
        part += float((acc.double() ** 2).sum())
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.tensor([part], dtype=torch.float64))
    norm = float(sum(float(x) for x in parts)) ** 0.5                # partial norms added in rank order on every rank
    coef = min(1.0, max_norm / (norm + 1e-6))
    p_sh = p0.clone()
    for s in range(3):
        p_new, _, _ = adamw(p0[lo[s]:hi[s]], reduced[s], m0[lo[s]:hi[s]], v0[lo[s]:hi[s]], coef)
        p_sh[lo[s]:hi[s]] = p_new
    for r in range(world):                                           # the all-gather (gather_master)
        l2, h2 = (ctypes.c_longlong * 3)(), (ctypes.c_longlong * 3)()
        lib.fnd_dp_shard_ranges(h, r, world, l2, h2)
        for s in range(3):
            if h2[s] > l2[s]:
                dist.broadcast(p_sh[l2[s]:h2[s]], src=r)
    # ---- replicated path: all-reduce, then the same AdamW everywhere ----
    total = torch.zeros(n)
    for r in range(world):
        total = total + all_g[r]
    norm_full = float((total.double() ** 2).sum()) ** 0.5
    p_full, _, _ = adamw(p0, total, m0, v0, min(1.0, max_norm / (norm_full + 1e-6)))
    if rank == 0:
        ret["norm_err"] = abs(norm - norm_full) / norm_full
        ret["equal"] = bool(torch.equal(p_sh, p_full)) if abs(norm - norm_full) == 0 else float((p_sh - p_full).abs().max())
    lib.fnd_plan_destroy(h)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_equals_replicated_optimizer():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_zero1_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["norm_err"] < 1e-12, ret["norm_err"]
    assert ret["equal"] is True or ret["equal"] < 1e-9, ret["equal"]


# ------------------------------------------------------------------------------------------------------------
# Epoch bookkeeping under data parallelism with n and batch_size NOT divisible by the world size:
# ragged shards are always gathered, every rank ends up with the same rows, the same mean-of-batch-means loss
# (forensic_trainer.py:301,316) and therefore the same early-stopping / checkpoint decisions.
# ------------------------------------------------------------------------------------------------------------
def _ragged_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ultrafnd_git_b200.trainer import gather_ragged, mean_of_batch_means, shard_indices
    n, bs = 23, 5                                  # 5 batches: 5,5,5,5,3 -> shards 3/2, ..., 2/1
    g = torch.Generator().manual_seed(3)
    row_loss = torch.rand(n, generator=g)           # "loss of sample i" as every rank would compute it
    order = torch.randperm(n, generator=g)
    cap = (n + world - 1) // world + (n + bs - 1) // bs + 1
    loss_rows, ids, bid = torch.zeros(cap), torch.zeros(cap, dtype=torch.int64), torch.zeros(cap, dtype=torch.int64)
    done = nb = 0
    for bno, s0 in enumerate(range(0, n, bs)):
        gidx = order[s0:s0 + bs]
        nb += 1
        local = shard_indices(gidx, rank, world)
        k = local.numel()
        loss_rows[done:done + k] = row_loss[local]
        ids[done:done + k] = local
        bid[done:done + k] = bno
        done += k
    (lr, ii, bb), total = gather_ragged([loss_rows, ids, bid], done, world)
    mine = mean_of_batch_means(lr, bb, nb)
    want = float(torch.stack([row_loss[order[s0:s0 + bs]].double().mean() for s0 in range(0, n, bs)]).mean())
    ret[rank] = (total, sorted(ii.tolist()) == list(range(n)), mine, want, float(row_loss.mean()))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_epoch_gather_is_identical_on_every_rank():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ragged_worker, args=(2, port, ret), nprocs=2, join=True)
    (t0, ok0, m0, want, rowmean), (t1, ok1, m1, _, _) = ret[0], ret[1]
    assert t0 == t1 == 23 and ok0 and ok1
    assert m0 == m1, "ranks must agree bit-for-bit (same rows, same order)"
    assert abs(m0 - want) < 1e-6
    assert abs(want - rowmean) > 1e-4, "case must distinguish mean-of-batch-means from the per-row mean"
