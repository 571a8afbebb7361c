"""Sequence front-end (Tier B): token / frame / audio SEQUENCES -> the pooled (B, D) vectors CrossModalTransformer consumes.

SELF-ORACLE SCOPE — the reference has no counterpart (no sequence attention, LayerNorm or per-token
projection exists in /root/reference). BASELINE.json's north_star names this front-end, so it is built here and checked
against oracle/seq_oracle.py (plain PyTorch); nothing in this file claims reference parity. The one reference semantic
carried over is the masked mean-pool, src/core_blocks/text_blocks.py:81-86.

    per stream s   : X_s = LayerNorm(x_s W_s^T + b_s)                      tcgen05 GEMM (fnd_seq_linear) + row kernel
    per block (a,b): bidirectional multi-head co-attention (d_k = 64), both directions from the block's input states,
                     X_a <- LayerNorm(X_a + MHA(q = X_a, kv = X_b, key-padding mask of b)), likewise X_b
                     fused [Q|K|V] projection GEMM per side, flash-style tcgen05 attention (fnd_seq_coattn_forward),
                     out-projection GEMM with the residual added in its epilogue, LayerNorm row kernel
    pool + head    : masked mean over valid tokens, Linear(d -> D_out) sized for CrossModalTransformer's inputs

Parameter names equal oracle/seq_oracle.py's (``embed.<s>.*``, ``embed_ln.<s>.*``, ``blocks.<i>.<a|b>.{in_proj,out_proj,ln}.*``,
``head.<s>.*``) so a state_dict moves between the two. Forward only (inference / feature extraction): the fused backward
is not built yet, outputs carry no autograd graph. There is no CPU path: every op raises without CUDA.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import seq_ops as S

# FakeSV / SV-FEND-shaped default (public setup restated in SURVEY.md §8d): stream -> (D_in, D_out, Tier-A feature key)
FAKESV_STREAMS: Dict[str, Tuple[int, int, str]] = {
    "text": (768, 768, "text_features"), "frames": (4096, 512, "visual_features"),
    "audio": (128, 128, "audio_features"), "c3d": (4096, 256, "temporal_features")}
FAKESV_BLOCKS: Tuple[Tuple[str, str], ...] = (("text", "frames"), ("text", "audio"))


class _Side(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.in_proj = nn.Linear(d, 3 * d)       # rows [0,d) Wq, [d,2d) Wk, [2d,3d) Wv
        self.out_proj = nn.Linear(d, d)
        self.ln = nn.LayerNorm(d)


class _CoBlock(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.a = _Side(d)
        self.b = _Side(d)


class SequenceFrontEnd(nn.Module):
    def __init__(self, d_model: int = 512, heads: int = 8, streams: Optional[Dict[str, Tuple[int, int, str]]] = None,
                 blocks: Optional[Sequence[Tuple[str, str]]] = None, eps: float = 1e-5):
        super().__init__()
        if d_model != heads * 64:
            raise NotImplementedError("the sm_100a attention kernel is built for head dimension 64: d_model must be 64 * heads")
        if d_model > 2048:
            raise NotImplementedError("d_model <= 2048 (LayerNorm row kernel keeps the row in registers)")
        self.d_model, self.heads, self.eps = d_model, heads, eps
        self.streams = dict(FAKESV_STREAMS if streams is None else streams)
        self.block_pairs = tuple(FAKESV_BLOCKS if blocks is None else blocks)
        for a, b in self.block_pairs:
            if a not in self.streams or b not in self.streams or a == b:
                raise ValueError(f"block ({a}, {b}) must name two different streams")
        for name, (din, dout, _key) in self.streams.items():
            if din % 8 or dout % 8:
                raise NotImplementedError("stream widths must be multiples of 8")
        self.embed = nn.ModuleDict({n: nn.Linear(s[0], d_model) for n, s in self.streams.items()})
        self.embed_ln = nn.ModuleDict({n: nn.LayerNorm(d_model) for n in self.streams})
        self.blocks = nn.ModuleList([_CoBlock(d_model) for _ in self.block_pairs])
        self.head = nn.ModuleDict({n: nn.Linear(d_model, s[1]) for n, s in self.streams.items()})
        self._shadow: Dict[str, torch.Tensor] = {}
        self._shadow_version = None
        self._err: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ bf16 operand copies of the GEMM weights
    def _version(self):
        ps = list(self.parameters())
        return (sum(p._version for p in ps), tuple(p.data_ptr() for p in ps))

    def _weights(self) -> Dict[str, torch.Tensor]:
        v = self._version()
        if v != self._shadow_version:
            self._shadow = {k: p.detach().to(torch.bfloat16).contiguous() for k, p in self.named_parameters()
                            if k.endswith("weight") and p.dim() == 2}
            self._shadow_version = v
        return self._shadow

    def check_error(self) -> None:
        if self._err is not None:
            code = int(self._err.item())
            if code:
                self._err.zero_()
                raise RuntimeError(f"sequence front-end kernel reported device-side error code {code}")

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, batch: Dict[str, torch.Tensor], return_states: bool = False) -> Dict[str, torch.Tensor]:
        """batch[name]: (B, L, D_in) fp32 or bf16; batch[name + "_mask"]: (B, L) bool, True = valid token (optional).
        Returns {name: (B, D_out) fp32} (+ "pooled.<name>" / "state.<name>" with return_states)."""
        first = next(iter(self.streams))
        dev = batch[first].device
        if dev.type != "cuda":
            raise RuntimeError("ultrafnd_git_b200.SequenceFrontEnd runs on CUDA (sm_100a) only: no CPU fallback exists")
        if self.embed[first].weight.device != dev:
            raise RuntimeError("module and inputs are on different devices")
        W = self._weights()
        if self._err is None or self._err.device != dev:
            self._err = S.new_err_flag(dev)
        err = self._err
        d, H = self.d_model, self.heads
        X: Dict[str, torch.Tensor] = {}
        shape: Dict[str, Tuple[int, int]] = {}
        mask_u8: Dict[str, Optional[torch.Tensor]] = {}
        length: Dict[str, Optional[torch.Tensor]] = {}
        for name in self.streams:
            x = batch[name]
            B, L, _ = x.shape
            shape[name] = (B, L)
            m = batch.get(name + "_mask")
            if m is not None:
                m = m.to(dev).bool()
                mask_u8[name] = m.to(torch.uint8).contiguous()
                pos = torch.arange(1, L + 1, device=dev, dtype=torch.int32)
                length[name] = (m.to(torch.int32) * pos).amax(dim=1).to(torch.int32).contiguous()   # last valid index + 1
            else:
                mask_u8[name], length[name] = None, None
            xb = x.contiguous() if x.dtype == torch.bfloat16 else S.cast_bf16(x.to(torch.float32))
            y = S.linear(xb.view(B * L, -1), W[f"embed.{name}.weight"], self.embed[name].bias, err=err)
            X[name] = S.layernorm(y, self.embed_ln[name].weight, self.embed_ln[name].bias, self.eps)
        for i, (a, b) in enumerate(self.block_pairs):
            blk = self.blocks[i]
            (Ba, La), (Bb, Lb) = shape[a], shape[b]
            qkv_a = S.linear(X[a], W[f"blocks.{i}.a.in_proj.weight"], blk.a.in_proj.bias, err=err)      # [B*La, 3d]
            qkv_b = S.linear(X[b], W[f"blocks.{i}.b.in_proj.weight"], blk.b.in_proj.bias, err=err)
            att_a = S.coattn_forward(qkv_a, qkv_b, qkv_b, Ba, H, La, Lb, q_col0=0, k_col0=d, v_col0=2 * d,
                                     kv_len=length[b], kv_mask=mask_u8[b], err=err)
            att_b = S.coattn_forward(qkv_b, qkv_a, qkv_a, Bb, H, Lb, La, q_col0=0, k_col0=d, v_col0=2 * d,
                                     kv_len=length[a], kv_mask=mask_u8[a], err=err)
            ya = S.linear(att_a, W[f"blocks.{i}.a.out_proj.weight"], blk.a.out_proj.bias, resid=X[a], err=err)
            yb = S.linear(att_b, W[f"blocks.{i}.b.out_proj.weight"], blk.b.out_proj.bias, resid=X[b], err=err)
            X[a] = S.layernorm(ya, blk.a.ln.weight, blk.a.ln.bias, self.eps)
            X[b] = S.layernorm(yb, blk.b.ln.weight, blk.b.ln.bias, self.eps)
        out: Dict[str, torch.Tensor] = {}
        for name, (_din, dout, _key) in self.streams.items():
            B, L = shape[name]
            pooled, pooled_bf = S.masked_mean_pool(X[name], B, L, mask=mask_u8[name], length=length[name], want_bf16=True)
            y = torch.empty(B, dout, dtype=torch.float32, device=dev)
            S.linear(pooled_bf, W[f"head.{name}.weight"], self.head[name].bias, out_f32=y, want_bf16=False, err=err)
            out[name] = y
            if return_states:
                out["pooled." + name] = pooled
                out["state." + name] = X[name].view(B, L, d)
        return out

    @torch.no_grad()
    def forward_features(self, feats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Tier-A keyed view: ``feats[<feature key>]`` is a (B, L, D_in) sequence (+ optional ``<feature key>_mask``);
        returns the dict CrossModalTransformer.forward expects (2-D vectors under the same keys; other entries such as
        ``gnn_feat`` pass through)."""
        batch: Dict[str, torch.Tensor] = {}
        for name, (_din, _dout, key) in self.streams.items():
            batch[name] = feats[key]
            if feats.get(key + "_mask") is not None:
                batch[name + "_mask"] = feats[key + "_mask"]
        res = self.forward(batch)
        out = {k: v for k, v in feats.items() if not k.endswith("_mask")}
        for name, (_din, _dout, key) in self.streams.items():
            out[key] = res[name]
        return out

    def flops(self, lengths: Dict[str, int], batch: int) -> float:
        """Forward FLOPs of the GEMM-shaped work (embeddings, in/out projections, scores, PV, heads)."""
        d = self.d_model
        f = 0.0
        for n, (din, dout, _k) in self.streams.items():
            f += 2.0 * batch * lengths[n] * din * d + 2.0 * batch * d * dout
        for a, b in self.block_pairs:
            La, Lb = lengths[a], lengths[b]
            f += batch * 2.0 * (2.0 * (2.0 * La * d * d + 2.0 * Lb * d * d + 2.0 * La * Lb * d))
        return f
