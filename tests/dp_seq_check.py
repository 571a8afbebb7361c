#!/usr/bin/env python3
"""Multi-GPU parity of the sequence front-end's data-parallel training step (seqfront.SequenceTrainer). Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tests/dp_seq_check.py

Every rank trains STEPS steps on its own slice of a global batch (loss = mean over the LOCAL samples of a regression
loss; gradients summed by the bucketed NCCL all-reduce and divided by the world size inside the optimizer kernel). Rank 0
then trains a single-GPU replica from the same initial weights on the WHOLE batch (loss = mean over all samples — the same
objective) and compares the loss trajectory, the clipped gradient norm of the first step and every parameter. The
per-rank partial sums are added in a different order and the bf16 activations of a 2 x B/2 evaluation round differently
from a 1 x B one in a few places, so agreement is to the front-end's bf16 tolerance (2e-2, north_star), not bitwise.
SELF-ORACLE SCOPE: the reference has no such front-end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from oracle import seq_oracle as O
from ultrafnd_git_b200.seqfront import SequenceFrontEnd, SequenceTrainer

STEPS = 4
D_MODEL, HEADS, B = 128, 2, 8
LENGTHS = {"text": 40, "frames": 83, "audio": 50, "c3d": 83}


def run(fe, tr, data, target, steps):
    losses, norm0 = [], None
    for it in range(steps):
        tr.zero_grad()
        out = fe(data)
        loss = sum(((out[n] - target[n]) ** 2).mean() for n in O.FAKESV_STREAMS)
        loss.backward()
        tr.step()
        if it == 0:
            norm0 = tr.grad_norm()
        losses.append(float(loss.detach()))
    return losses, norm0


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    streams, blocks = O.FAKESV_STREAMS, O.FAKESV_BLOCKS
    p = O.init_params(streams, blocks, D_MODEL, seed=3)
    full = O.make_batch(streams, LENGTHS, B, seed=4)
    g = torch.Generator().manual_seed(5)
    tgt = {n: torch.randn(B, streams[n][1], generator=g) for n in streams}
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    data = {k: v[sl].cuda() for k, v in full.items()}
    target = {n: v[sl].cuda() for n, v in tgt.items()}
    fe = SequenceFrontEnd(D_MODEL, HEADS, streams, blocks).cuda(); fe.load_state_dict(p)
    tr = SequenceTrainer(fe, lr=2e-3, max_norm=1.0)
    losses, norm0 = run(fe, tr, data, target, STEPS)
    lt = torch.tensor(losses, device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)                                   # mean over ranks of the local means = the global mean loss
    lt /= world
    fe.check_error()
    # every rank must hold bit-identical parameters (same reduced gradients, same optimizer arithmetic)
    chk = tr.flat_w.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(bool(torch.equal(allc[0], c)) for c in allc)
    ok = True
    if rank == 0:
        ref = SequenceFrontEnd(D_MODEL, HEADS, streams, blocks).cuda(); ref.load_state_dict(p)
        rt = SequenceTrainer(ref, lr=2e-3, max_norm=1.0, world=1)      # single-GPU replica: no exchange
        rl, rn = run(ref, rt, {k: v.cuda() for k, v in full.items()}, {n: v.cuda() for n, v in tgt.items()}, STEPS)
        e_loss = max(abs(a - b) / abs(b) for a, b in zip(lt.tolist(), rl))
        e_norm = abs(norm0 - rn) / rn
        w0 = torch.cat([p[k].reshape(-1) for bk in fe.grad_order() for k in bk]).cuda().double()      # initial weights, flat order
        upd, upd_ref = tr.flat_w.double() - w0, rt.flat_w.double() - w0
        e_par = float((upd - upd_ref).norm() / upd_ref.norm())
        print(f"[dp_seq world={world}] losses {['%.5f' % x for x in lt.tolist()]} vs single GPU {['%.5f' % x for x in rl]}")
        print(f"[dp_seq world={world}] loss rel-err {e_loss:.2e}, first-step gradient norm {norm0:.5f} vs {rn:.5f} ({e_norm:.2e}), "
              f"parameter UPDATE rel-L2 error after {STEPS} steps {e_par:.2e}, replicas identical: {same}")
        ok = same and e_loss < 2e-2 and e_norm < 2e-2 and e_par < 1e-1
        print("dp_seq_check", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
