#!/usr/bin/env python3
"""Event stamps of the attention forward pipeline (probe build hook fnd_seq_debug_attn_stamps): per CTA and key-block step,
clock64 at: softmax tile t sees S (0+4t), S in registers (1+4t), exp2 section entered (2+4t), P published (3+4t);
MMA warp sees P_t (8+4t), P V issued (9+4t), next Q K^T issued (10+4t)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ultrafnd_git_b200 import seq_ops as S, _lib

B, Lq, Lk, H = 32, 1024, 512, 16
d = H * 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B * Lq, 3 * d, device=dev, generator=g).bfloat16()
k = torch.randn(B * Lk, 3 * d, device=dev, generator=g).bfloat16()
o = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
lib = _lib.load()
for _ in range(3):
    S.coattn_forward(q, k, k, B, H, Lq, Lk, 0, d, 2 * d, out=o)
st = torch.zeros(148 * 32 * 16, dtype=torch.int64, device=dev)
lib.fnd_seq_debug_attn_stamps(st.data_ptr())
S.coattn_forward(q, k, k, B, H, Lq, Lk, 0, d, 2 * d, out=o)
torch.cuda.synchronize()
lib.fnd_seq_debug_attn_stamps(None)
v = st.view(148, 32, 16).cpu().double()
names = {0: "T0 sees S", 1: "T0 S in regs", 2: "T0 exp enter", 3: "T0 P published", 4: "T1 sees S", 5: "T1 S in regs", 6: "T1 exp enter",
         7: "T1 P published", 8: "MMA0 sees P0", 9: "MMA0 PV0 issued", 10: "MMA0 next QK0 issued", 12: "MMA1 sees P1", 13: "MMA1 PV1 issued",
         14: "MMA1 next QK1 issued"}
for cta in (0, 77):
    base = v[cta, 8, 0]
    print(f"CTA {cta}: steps 8..13, cycles relative to 'T0 sees S' of step 8")
    for step in range(8, 14):
        evs = sorted(((float(v[cta, step, e] - base), names[e]) for e in names if v[cta, step, e] > 0))
        print(f"  step {step}: " + " | ".join(f"{n} {t:.0f}" for t, n in evs))
# averages over CTAs and steps 6..25
def delta(a, b, sa=0, sb=0):
    x = v[:, 6 + sb:26 + sb, b] - v[:, 6 + sa:26 + sa, a]
    return float(x.mean())
print("mean deltas (cycles), steps 6..25, all CTAs:")
for t in (0, 1):
    o4 = 4 * t
    print(f" tile {t}: sees S -> S in regs {delta(0 + o4, 1 + o4):.0f} | -> exp2 section entered {delta(1 + o4, 2 + o4):.0f} | -> P published "
          f"{delta(2 + o4, 3 + o4):.0f} | P published -> MMA warp sees it {delta(3 + o4, 8 + o4):.0f} | -> P V issued {delta(8 + o4, 9 + o4):.0f} | "
          f"P published -> softmax sees the next S {delta(3 + o4, 0 + o4, 0, 1):.0f} | step period {delta(0 + o4, 0 + o4, 0, 1):.0f}")
