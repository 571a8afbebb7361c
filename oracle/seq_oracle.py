"""CPU self-oracle for the sequence front-end (Tier B).  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

**SELF-ORACLE — PARITY UNPINNED BY THE REFERENCE.**  The reference checkout contains no sequence-level co-attention,
no LayerNorm and no per-token projection (grep for MultiheadAttention / scaled_dot_product / LayerNorm
over /root/reference gives 0 hits). BASELINE.json's north_star nevertheless names that front-end (token / frame / audio
sequences -> per-modality projection + LayerNorm -> bidirectional multi-head co-attention with key-padding masks ->
masked mean-pool -> the (B, D) vectors the reference's CrossModalTransformer consumes), so this file restates the
*textbook* operators in plain PyTorch fp32/fp64 and the CUDA path is checked against it. No number produced with this
file is a statement about the reference. The single reference anchor is the masked mean-pool semantics:
``sum(x * m) / clamp_min(sum(m), 1e-6)`` — src/core_blocks/text_blocks.py:81-86.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this module.

Model (all shapes from the public FakeSV / SV-FEND setup restated in SURVEY.md §8d, not from /root/reference):
  streams   : name -> (D_in, D_out); each stream is a padded batch x[B, L, D_in] with a validity mask m[B, L]
  embed     : X_s = LayerNorm_s(x_s W_s^T + b_s)                                        (B, L_s, d)
  blocks    : list of (a, b): bidirectional co-attention between streams a and b, both directions computed from the
              block's INPUT states:
                 A = MHA(q = X_a, k = v = X_b, key_padding_mask = ~m_b);  B = MHA(q = X_b, k = v = X_a, ~m_a)
                 X_a <- LayerNorm(X_a + A);  X_b <- LayerNorm(X_b + B)                    (post-LN residual)
              MHA(q, kv) = concat_h softmax(Q_h K_h^T / sqrt(d_k) + mask) V_h  W_o^T + b_o with Q = q Wq^T + bq etc.
              A query row whose keys are ALL masked gets a zero attention output (torch would give NaN).
  pool      : p_s = sum_l X_s[l] m[l] / clamp_min(sum_l m[l], 1e-6)                      (text_blocks.py:81-86)
  heads     : y_s = p_s Wh_s^T + bh_s   (D_out = the width CrossModalTransformer expects for that modality)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# FakeSV-shaped default: stream -> (D_in, D_out, feature key of the Tier-A model it feeds)
FAKESV_STREAMS = {"text": (768, 768, "text_features"), "frames": (4096, 512, "visual_features"),
                  "audio": (128, 128, "audio_features"), "c3d": (4096, 256, "temporal_features")}
FAKESV_BLOCKS = (("text", "frames"), ("text", "audio"))
LN_EPS = 1e-5


def param_shapes(streams: Dict[str, Tuple], blocks: Sequence[Tuple[str, str]], d_model: int) -> Dict[str, Tuple[int, ...]]:
    d = d_model
    s: Dict[str, Tuple[int, ...]] = {}
    for name, spec in streams.items():
        s[f"embed.{name}.weight"] = (d, spec[0]); s[f"embed.{name}.bias"] = (d,)
        s[f"embed_ln.{name}.weight"] = (d,); s[f"embed_ln.{name}.bias"] = (d,)
    for i, (a, b) in enumerate(blocks):
        for side in ("a", "b"):
            # in_proj: rows [0,d) = Wq (this side as QUERY), [d,2d) = Wk, [2d,3d) = Wv (this side as KEY/VALUE source)
            s[f"blocks.{i}.{side}.in_proj.weight"] = (3 * d, d); s[f"blocks.{i}.{side}.in_proj.bias"] = (3 * d,)
            s[f"blocks.{i}.{side}.out_proj.weight"] = (d, d); s[f"blocks.{i}.{side}.out_proj.bias"] = (d,)
            s[f"blocks.{i}.{side}.ln.weight"] = (d,); s[f"blocks.{i}.{side}.ln.bias"] = (d,)
    for name, spec in streams.items():
        s[f"head.{name}.weight"] = (spec[1], d); s[f"head.{name}.bias"] = (spec[1],)
    return s


def init_params(streams, blocks, d_model: int, seed: int = 7, dtype=torch.float32) -> Params:
    """nn.Linear-style uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights, LayerNorm gains perturbed around 1 so that the
    affine part is exercised."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for k, shp in param_shapes(streams, blocks, d_model).items():
        if ".ln." in k or k.startswith("embed_ln."):
            base = 1.0 if k.endswith("weight") else 0.0
            p[k] = (base + 0.1 * torch.randn(shp, generator=g)).to(dtype)
        else:
            fan_in = shp[-1] if len(shp) == 2 else None
            if fan_in is None:
                fan_in = param_shapes(streams, blocks, d_model)[k.replace(".bias", ".weight")][-1]
            bound = 1.0 / math.sqrt(fan_in)
            p[k] = ((torch.rand(shp, generator=g) * 2 - 1) * bound).to(dtype)
    return p


def make_batch(streams, lengths: Dict[str, int], batch: int, seed: int = 1234, full: bool = False,
               dtype=torch.float32) -> Dict[str, Tensor]:
    """randn features with random valid lengths in [1, L] (prefix masks = key-padding masks); ``full`` = no padding.
    One sample per batch additionally gets a scattered (non-prefix) mask on the first stream to exercise general masks."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for i, (name, spec) in enumerate(streams.items()):
        L = lengths[name]
        out[name] = torch.randn(batch, L, spec[0], generator=g).to(dtype)
        if full:
            m = torch.ones(batch, L, dtype=torch.bool)
        else:
            n = torch.randint(1, L + 1, (batch,), generator=g)
            n[0] = L
            m = torch.arange(L)[None, :] < n[:, None]
            if i == 0 and batch > 1 and L >= 8:
                m[1] = torch.rand(L, generator=g) < 0.7
                m[1, 0] = True
        out[name + "_mask"] = m
    return out


def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def mha(xq: Tensor, xkv: Tensor, kv_mask: Tensor, p: Params, q_side: str, kv_side: str, heads: int) -> Tensor:
    """softmax(QK^T / sqrt(d_k) + key_padding_mask) V per head, then the QUERY side's out_proj. Q uses rows [0,d) of the
    query side's in_proj; K / V use rows [d,2d) / [2d,3d) of the KEY side's in_proj."""
    B, Lq, d = xq.shape
    Lk = xkv.shape[1]
    dk = d // heads
    Wq, bq = p[f"{q_side}.in_proj.weight"][:d], p[f"{q_side}.in_proj.bias"][:d]
    Wk, bk = p[f"{kv_side}.in_proj.weight"][d:2 * d], p[f"{kv_side}.in_proj.bias"][d:2 * d]
    Wv, bv = p[f"{kv_side}.in_proj.weight"][2 * d:], p[f"{kv_side}.in_proj.bias"][2 * d:]
    q = F.linear(xq, Wq, bq).view(B, Lq, heads, dk).transpose(1, 2)
    k = F.linear(xkv, Wk, bk).view(B, Lk, heads, dk).transpose(1, 2)
    v = F.linear(xkv, Wv, bv).view(B, Lk, heads, dk).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
    s = s.masked_fill(~kv_mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    a = torch.nan_to_num(a, nan=0.0)                 # no valid key at all -> zero output
    o = (a @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, p[f"{q_side}.out_proj.weight"], p[f"{q_side}.out_proj.bias"])


def masked_mean(x: Tensor, m: Tensor) -> Tensor:
    """src/core_blocks/text_blocks.py:81-86."""
    mf = m.unsqueeze(-1).to(x.dtype)
    return (x * mf).sum(dim=1) / mf.sum(dim=1).clamp_min(1e-6)


def forward(p: Params, batch: Dict[str, Tensor], streams, blocks, heads: int,
            return_states: bool = False) -> Dict[str, Tensor]:
    X: Dict[str, Tensor] = {}
    for name in streams:
        X[name] = layer_norm(F.linear(batch[name], p[f"embed.{name}.weight"], p[f"embed.{name}.bias"]),
                             p[f"embed_ln.{name}.weight"], p[f"embed_ln.{name}.bias"])
    for i, (a, b) in enumerate(blocks):
        pa, pb = f"blocks.{i}.a", f"blocks.{i}.b"
        A = mha(X[a], X[b], batch[b + "_mask"], p, pa, pb, heads)
        Bo = mha(X[b], X[a], batch[a + "_mask"], p, pb, pa, heads)
        X[a] = layer_norm(X[a] + A, p[f"{pa}.ln.weight"], p[f"{pa}.ln.bias"])
        X[b] = layer_norm(X[b] + Bo, p[f"{pb}.ln.weight"], p[f"{pb}.ln.bias"])
    out: Dict[str, Tensor] = {}
    for name in streams:
        pooled = masked_mean(X[name], batch[name + "_mask"])
        out[name] = F.linear(pooled, p[f"head.{name}.weight"], p[f"head.{name}.bias"])
        if return_states:
            out["pooled." + name] = pooled
            out["state." + name] = X[name]
    return out


def attention_only(q: Tensor, k: Tensor, v: Tensor, kv_mask: Optional[Tensor], heads: int) -> Tuple[Tensor, Tensor]:
    """The attention core alone on already-projected [B, L, d] tensors: returns (O [B, Lq, d], LSE [B, heads, Lq])."""
    B, Lq, d = q.shape
    Lk = k.shape[1]
    dk = d // heads
    qh = q.view(B, Lq, heads, dk).transpose(1, 2)
    kh = k.view(B, Lk, heads, dk).transpose(1, 2)
    vh = v.view(B, Lk, heads, dk).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) / math.sqrt(dk)
    if kv_mask is not None:
        s = s.masked_fill(~kv_mask[:, None, None, :], float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    a = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    return (a @ vh).transpose(1, 2).reshape(B, Lq, d), lse


def rel_err(a: Tensor, b: Tensor) -> float:
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def layer_flops(batch: int, La: int, Lb: int, d: int) -> float:
    """forward FLOPs of ONE bidirectional co-attention block (in/out projections + scores + PV),
    2 * [2 * (2 La d^2 + 2 Lb d^2 + 2 La Lb d)] per sample."""
    return batch * 2.0 * (2.0 * (2.0 * La * d * d + 2.0 * Lb * d * d + 2.0 * La * Lb * d))
