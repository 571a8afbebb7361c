"""fnd_train_step vs fnd_train_step_overlap at the bench configuration (batch 128, bf16, CUDA graph, L2 flushed):
step time of both, and a bit-identity check of the parameters / loss after a few steps with dropout ON (the two entry
points run the same tiles and write the same norm slots, so nothing may differ).

    python tools/overlap_probe.py [batch] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ultrafnd_git_b200.fused import FusedStep  # noqa: E402
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier  # noqa: E402
import bench  # noqa: E402


def build(B, early, precision="bf16"):
    torch.manual_seed(0)
    f = CrossModalTransformer(precision=precision)
    c = DeepTruthClassifier(precision=precision)
    f.train(); c.train()
    os.environ["FND_WG_EARLY"] = "1" if early else "0"
    step = FusedStep(f, c, B, precision=precision, use_graph=True)
    return f, c, step


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    batch = bench.synth_batch(B, 5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for early in (False, True, False, True):
        f, c, step = build(B, early)
        step.load_batch({k: v.cuda() for k, v in batch.items()})
        for _ in range(5):
            step.train_step()
        torch.cuda.synchronize()
        step.plan.check_error()
        if early not in res:
            res[early] = (step.plan.state()["loss"], {k: v.detach().clone() for k, v in list(f.state_dict().items()) + list(c.state_dict().items())})
        ts = []
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step.train_step(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print(f"early={int(early)}  median {ts[len(ts) // 2]:.1f} us  p10 {ts[len(ts) // 10]:.1f}  min {ts[0]:.1f}", flush=True)
        step.plan.check_error()
    l0, p0 = res[False]
    l1, p1 = res[True]
    bad = [k for k in p0 if not torch.equal(p0[k], p1[k])]
    print(f"loss after 5 steps: {l0!r} vs {l1!r}; tensors differing: {len(bad)} of {len(p0)}", bad[:5])
    assert l0 == l1 and not bad, "fnd_train_step_overlap is not bit-identical to fnd_train_step"
    print("overlap_probe OK")


if __name__ == "__main__":
    main()
