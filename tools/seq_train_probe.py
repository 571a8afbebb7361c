#!/usr/bin/env python3
"""One training step (forward with activations kept + fused backward) of the sequence front-end at the stress shape; wrapped by
ncu for the per-launch time list in profiles/ (usage: python tools/seq_train_probe.py [B] [steps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench_stress as BS
from ultrafnd_git_b200.seqfront import SequenceFrontEnd


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda")
    torch.manual_seed(42)
    fe = SequenceFrontEnd(BS.D_MODEL, BS.HEADS, BS.STREAMS, BS.BLOCKS).to(dev)
    g = torch.Generator(device="cuda").manual_seed(1)
    batch = {"text": torch.randn(B, BS.LT, 768, device=dev, generator=g), "frames": torch.randn(B, BS.LF, 4096, device=dev, generator=g)}
    w = {n: torch.randn(B, s[1], device=dev, generator=g) for n, s in BS.STREAMS.items()}
    for i in range(steps):
        for p in fe.parameters():
            p.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = fe(batch)
        sum((o[n] * w[n]).sum() for n in BS.STREAMS).backward()
        b.record()
        torch.cuda.synchronize()
        print(f"step {i}: {a.elapsed_time(b):.3f} ms")
    fe.check_error()


if __name__ == "__main__":
    main()
