"""CPU oracle for the Ultrafnd fusion hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module, and only as the checker or as the timed CPU baseline. Nothing under
``ultrafnd_git_b200/`` imports it; the product path raises when the CUDA library is missing.

What it is: a functional, from-scratch restatement (plain torch CPU ops on an explicit ``{name: tensor}``
parameter dict keyed like the reference's ``state_dict``) of the reference's
``CrossModalTransformer.forward`` → ``DeepTruthClassifier.forward`` → ``F.cross_entropy`` →
``clip_grad_norm_`` → ``AdamW`` step. Every function cites the reference lines it follows. All arithmetic
is floating point (fp32 by default, fp64 on request for a tighter truth when judging bf16).

Pinning: the reference ships NO golden vectors or numeric tests. The oracle is pinned
against outputs of the reference itself, produced in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this module against those fixtures on every CPU run.

Paths are relative to the reference checkout (Nuralamsiddik16/Ultrafnd_git).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

HIDDEN = 512
IN_DIMS = {"text": 768, "audio": 128, "visual": 512, "temporal": 256, "gnn": 128}
FEAT_KEYS = ("text_features", "audio_features", "visual_features", "temporal_features", "gnn_feat")
ATTN_BLOCKS = ("attn_tv", "attn_ta", "attn_vu")
NODE_TREES, NODE_DEPTH, NODE_TAU = 6, 4, 10.0


# --------------------------------------------------------------------------------------------------
# Parameter inventory and initialisation
# --------------------------------------------------------------------------------------------------
def fusion_param_shapes(hidden: int = HIDDEN, use_gnn: bool = True, gnn_dim: int = 128) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys/shapes of CrossModalTransformer, in registration order
    (src/models/fusion/cross_modal_transformer.py:96-130; semantic.* from src/models/semantic_forgery.py)."""
    H = hidden
    s: Dict[str, Tuple[int, ...]] = {}
    for name, d in (("text_proj", 768), ("audio_proj", 128), ("visual_proj", 512), ("temporal_proj", 256)):
        s[f"{name}.weight"] = (H, d)
        s[f"{name}.bias"] = (H,)
    if use_gnn:
        s["gnn_proj.weight"] = (H, gnn_dim)
        s["gnn_proj.bias"] = (H,)
    for name in ("semantic.text_proj.0", "semantic.vision_proj.0"):      # constructed, never used in forward
        s[f"{name}.weight"] = (512, 512)
        s[f"{name}.bias"] = (512,)
    for blk in ATTN_BLOCKS:
        for lin in ("q", "k", "v"):
            s[f"{blk}.{lin}.weight"] = (H, H)
            s[f"{blk}.{lin}.bias"] = (H,)
        s[f"{blk}.evidence_proj.0.weight"] = (H, 3)
        s[f"{blk}.evidence_proj.0.bias"] = (H,)
        s[f"{blk}.evidence_proj.2.weight"] = (1, H)
        s[f"{blk}.evidence_proj.2.bias"] = (1,)
    fused_dim = (4 + 8 + 3 + (1 if use_gnn else 0)) * H
    s["fuse_mlp.0.weight"] = (2 * H, fused_dim)
    s["fuse_mlp.0.bias"] = (2 * H,)
    s["fuse_mlp.3.weight"] = (H, 2 * H)
    s["fuse_mlp.3.bias"] = (H,)
    s["classifier.weight"] = (2, H)
    s["classifier.bias"] = (2,)
    return s


def classifier_param_shapes(hidden: int = HIDDEN, in_dim: int = 512, aux_dim: int = 2, trees: int = NODE_TREES,
                            depth: int = NODE_DEPTH, classes: int = 2) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys/shapes of DeepTruthClassifier (src/models/fusion/deep_truth_classifier.py:104-140)."""
    s: Dict[str, Tuple[int, ...]] = {"temperature": ()}
    s["pre.0.weight"] = (hidden, in_dim + aux_dim)
    s["pre.0.bias"] = (hidden,)
    s["pre.3.weight"] = (hidden, hidden)
    s["pre.3.bias"] = (hidden,)
    for t in range(trees):
        s[f"node.trees.{t}.tau"] = ()
        s[f"node.trees.{t}.leaf_logits"] = (1 << depth, classes)
        for k in range(depth):
            s[f"node.trees.{t}.gates.{k}"] = (hidden,)
        for k in range(depth):
            s[f"node.trees.{t}.thresh.{k}"] = (1,)
    s["bypass.weight"] = (classes, hidden)
    s["bypass.bias"] = (classes,)
    return s


def _linear_default_init(w: Tensor, b: Optional[Tensor], gen: torch.Generator) -> None:
    # nn.Linear.reset_parameters: kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    bound = 1.0 / math.sqrt(w.shape[1])
    w.uniform_(-bound, bound, generator=gen)
    if b is not None:
        b.uniform_(-bound, bound, generator=gen)


def init_params(seed: int = 42, dtype: torch.dtype = torch.float32) -> Tuple[Params, Params]:
    """Random-init weights with the reference's DISTRIBUTIONS (not its RNG stream): default nn.Linear init for
    fusion (cross_modal_transformer.py:96-130), xavier_uniform + zero bias for pre/bypass
    (deep_truth_classifier.py:18-21,129-130,138), zeros for gates/thresh/leaf tables (:44-50), tau=10, T=1."""
    gen = torch.Generator().manual_seed(seed)
    fus: Params = {k: torch.zeros(s, dtype=dtype) for k, s in fusion_param_shapes().items()}
    for k in list(fus):
        if k.endswith(".weight"):
            _linear_default_init(fus[k], fus[k[:-6] + "bias"], gen)
    clf: Params = {k: torch.zeros(s, dtype=dtype) for k, s in classifier_param_shapes().items()}
    clf["temperature"].fill_(1.0)
    for k in clf:
        if k.endswith(".tau"):
            clf[k].fill_(NODE_TAU)
    for k in ("pre.0.weight", "pre.3.weight", "bypass.weight"):
        fan_out, fan_in = clf[k].shape
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        clf[k].uniform_(-bound, bound, generator=gen)
    return fus, clf


def perturb_node_head(clf: Params, seed: int = 7, std: float = 0.02) -> None:
    """'Trained-like' state: gates/thresholds/leaf tables are zero at init, which makes every tree identical and
    the NODE head inert; add N(0, std) so parity tests exercise it."""
    gen = torch.Generator().manual_seed(seed)
    for k, v in clf.items():
        if ".gates." in k or ".thresh." in k or k.endswith("leaf_logits"):
            v.add_(torch.randn(v.shape, generator=gen, dtype=v.dtype) * std)


def trainable_keys() -> Tuple[List[str], List[str]]:
    """Parameters that receive a gradient in the reference training step (forensic_trainer.py:286-298):
    everything except semantic.*, fusion.classifier.*, temperature and tau."""
    fk = [k for k in fusion_param_shapes() if not k.startswith("semantic.") and not k.startswith("classifier.")]
    ck = [k for k in classifier_param_shapes() if k != "temperature" and not k.endswith(".tau")]
    return fk, ck


# --------------------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------------------
def make_batch(batch: int, seed: int = 1234, dist: str = "smoke", dtype: torch.dtype = torch.float32) -> Dict[str, Tensor]:
    """D1 'smoke': randn features, rand aux, randint labels (scripts/smoke_test_v2.py:43-45,55).
    D2 'cache': L2-normalised non-negative sparse rows for text/audio/visual (hash bag-of-words,
    core_blocks/text_blocks.py:19-27,128), small dense temporal, aux in [0,1]^2, small dense gnn."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    dims = [("text_features", 768), ("audio_features", 128), ("visual_features", 512), ("temporal_features", 256),
            ("gnn_feat", 128)]
    if dist == "smoke":
        for k, d in dims:
            out[k] = torch.randn(batch, d, generator=g, dtype=dtype)
    elif dist == "cache":
        for k, d in dims[:3]:
            x = torch.rand(batch, d, generator=g, dtype=dtype)
            x = torch.where(torch.rand(batch, d, generator=g) < 0.08, x, torch.zeros_like(x))
            out[k] = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-9)
        out["temporal_features"] = 0.05 * torch.randn(batch, 256, generator=g, dtype=dtype)
        out["gnn_feat"] = 0.1 * torch.randn(batch, 128, generator=g, dtype=dtype)
    else:
        raise ValueError(dist)
    out["aux"] = torch.rand(batch, 2, generator=g, dtype=dtype)
    out["label"] = torch.randint(0, 2, (batch,), generator=g)
    return out


# --------------------------------------------------------------------------------------------------
# Forward
# --------------------------------------------------------------------------------------------------
def _lin(p: Params, name: str, x: Tensor) -> Tensor:
    return x @ p[f"{name}.weight"].t() + p[f"{name}.bias"]


def _drop(x: Tensor, p_drop: float, masks: Optional[Dict[str, Tensor]], key: str) -> Tensor:
    """Inverted dropout with an EXPLICIT keep-multiplier (0 or 1/(1-p)) so the CUDA path's Philox masks can be
    exported and replayed here bit-for-bit; masks=None means eval mode / dropout off."""
    if masks is None or p_drop <= 0.0:
        return x
    return x * masks[key].to(x.dtype)


def evidence_scalars(t: Tensor, v: Tensor, u: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """cross_modal_transformer.py:153-164 — computed under no_grad in the reference."""
    with torch.no_grad():
        def cos01(a: Tensor, b: Tensor) -> Tensor:
            an = a / a.norm(dim=-1, keepdim=True).clamp_min(1e-12)     # F.normalize eps
            bn = b / b.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            return 0.5 * ((an * bn).sum(-1, keepdim=True).clamp(-1, 1) + 1.0)
        semantic_conflict = 1.0 - cos01(t, v)
        emo = t.abs().mean(dim=-1, keepdim=True).tanh()
        delay = 1.0 - cos01(t, u)
    return semantic_conflict, emo, delay


def co_attention(p: Params, blk: str, x: Tensor, y: Tensor, evidence: Tensor) -> Tensor:
    """ForensicCoAttention.forward (cross_modal_transformer.py:39-55): vector-level, sigmoid score, evidence gate."""
    H = x.shape[-1]
    q, k, v = _lin(p, f"{blk}.q", x), _lin(p, f"{blk}.k", y), _lin(p, f"{blk}.v", y)
    attn = torch.sigmoid((q * k).sum(-1, keepdim=True) / (H ** 0.5))
    hid = F.gelu(_lin(p, f"{blk}.evidence_proj.0", evidence))
    gate = torch.sigmoid(_lin(p, f"{blk}.evidence_proj.2", hid))
    return gate * (attn * v) + (1.0 - gate) * (0.5 * (x + y))


def fusion_forward(p: Params, feats: Dict[str, Tensor], dropout: float = 0.1,
                   masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """CrossModalTransformer.forward (cross_modal_transformer.py:134-210). Returns fused, logits, forensic and
    the intermediate ``fused_cat`` (for layer-level parity tests)."""
    t = _lin(p, "text_proj", feats["text_features"])
    a = _lin(p, "audio_proj", feats["audio_features"])
    v = _lin(p, "visual_proj", feats["visual_features"])
    u = _lin(p, "temporal_proj", feats["temporal_features"])
    sc, emo, delay = evidence_scalars(t, v, u)
    z = torch.zeros_like(emo)
    tv = co_attention(p, "attn_tv", t, v, torch.cat([sc, emo, z], -1))
    ta = co_attention(p, "attn_ta", t, a, torch.cat([emo, z, z], -1))
    vu = co_attention(p, "attn_vu", v, u, torch.cat([delay, z, z], -1))
    pairs = [t + a, t * a, (t - a).abs(), t + v, t * v, (t - v).abs(), t + u, v + u]
    parts = [t, a, v, u] + pairs + [tv, ta, vu]
    if feats.get("gnn_feat") is not None:
        parts.append(_lin(p, "gnn_proj", feats["gnn_feat"]))
    cat = torch.cat(parts, dim=-1)
    h1 = _drop(F.gelu(_lin(p, "fuse_mlp.0", cat)), dropout, masks, "fuse0")
    fused = _drop(F.gelu(_lin(p, "fuse_mlp.3", h1)), dropout, masks, "fuse1")
    logits = _lin(p, "classifier", fused)
    return {"fused": fused, "logits": logits, "fused_cat": cat,
            "forensic": {"emotion_intensity": emo.squeeze(-1), "semantic_conflict": sc.squeeze(-1),
                         "temporal_delay": delay.squeeze(-1)}}


def node_ensemble(p: Params, h: Tensor, masks: Optional[Dict[str, Tensor]] = None,
                  trees: int = NODE_TREES, depth: int = NODE_DEPTH) -> Tensor:
    """NODEEnsemble / _ObliviousTree (deep_truth_classifier.py:54-74,88-90): soft oblivious trees, leaf index
    doubles per depth as cat([p*(1-s), p*s]); Dropout(0.3) on each tree's logits; mean over trees."""
    outs = []
    for i in range(trees):
        tau = p[f"node.trees.{i}.tau"]
        probs = h.new_ones((h.shape[0], 1))
        for k in range(depth):
            alpha = torch.softmax(p[f"node.trees.{i}.gates.{k}"], dim=0)
            feat = (h * alpha).sum(-1, keepdim=True)
            s = torch.sigmoid(tau * (feat - p[f"node.trees.{i}.thresh.{k}"]))
            probs = torch.cat([probs * (1.0 - s), probs * s], dim=1)
        tl = probs @ p[f"node.trees.{i}.leaf_logits"]
        if masks is not None:
            tl = tl * masks["tree"][:, i, :].to(tl.dtype)
        outs.append(tl)
    return torch.stack(outs, 0).mean(0)


def classifier_forward(p: Params, fused: Tensor, aux: Optional[Tensor], dropout: float = 0.1,
                       masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """DeepTruthClassifier.forward (deep_truth_classifier.py:148-171)."""
    x = torch.cat([fused, aux], -1) if aux is not None else fused
    h = _drop(F.gelu(_lin(p, "pre.0", x)), dropout, masks, "pre0")
    h = _drop(F.gelu(_lin(p, "pre.3", h)), dropout, masks, "pre1")
    logits = node_ensemble(p, h, masks) + _lin(p, "bypass", h)
    temp = torch.clamp(p["temperature"], min=0.5, max=5.0)
    return {"logits": logits, "probs": torch.softmax(logits / temp, dim=-1), "temperature": temp, "h": h}


def model_forward(fus: Params, clf: Params, batch: Dict[str, Tensor], dropout: float = 0.1,
                  masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """ForensicTrainer._forward_batch (src/training/forensic_trainer.py:254-271) + the loss at :287."""
    fo = fusion_forward(fus, batch, dropout, masks)
    co = classifier_forward(clf, fo["fused"], batch["aux"], dropout, masks)
    out = {"fused": fo["fused"], "fusion_logits": fo["logits"], "fused_cat": fo["fused_cat"],
           "forensic": fo["forensic"], "logits": co["logits"], "probs": co["probs"], "h": co["h"]}
    if "label" in batch:
        out["loss"] = F.cross_entropy(co["logits"], batch["label"])
    return out


# --------------------------------------------------------------------------------------------------
# Training step
# --------------------------------------------------------------------------------------------------
def loss_and_grads(fus: Params, clf: Params, batch: Dict[str, Tensor], dropout: float = 0.1,
                   masks: Optional[Dict[str, Tensor]] = None) -> Tuple[Dict[str, Tensor], Params, Params]:
    """loss.backward() of forensic_trainer.py:287-291 via autograd on leaf copies of the trainable parameters."""
    fk, ck = trainable_keys()
    fl = {k: (v.detach().clone().requires_grad_(k in fk)) for k, v in fus.items()}
    cl = {k: (v.detach().clone().requires_grad_(k in ck)) for k, v in clf.items()}
    out = model_forward(fl, cl, batch, dropout, masks)
    out["loss"].backward()
    gf = {k: fl[k].grad for k in fk}
    gc = {k: cl[k].grad for k in ck}
    return {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in out.items()}, gf, gc


class AdamWState:
    """torch.optim.AdamW semantics (decoupled decay, bias correction, eps outside the sqrt of the corrected v) as
    used at forensic_trainer.py:173-177,298, plus clip_grad_norm_(max_norm=5.0) of :292-297 with torch's
    coefficient min(1, max_norm / (norm + 1e-6))."""

    def __init__(self, lr: float = 2e-4, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_clip: float = 5.0):
        self.lr, self.wd, self.betas, self.eps, self.clip = lr, weight_decay, betas, eps, grad_clip
        self.t = 0
        self.m: Dict[str, Tensor] = {}
        self.v: Dict[str, Tensor] = {}

    def step(self, params: Params, grads: Params) -> float:
        total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
        coef = min(1.0, self.clip / (total + 1e-6)) if self.clip and self.clip > 0 else 1.0
        self.t += 1
        b1, b2 = self.betas
        for k, g in grads.items():
            g = g * coef
            if k not in self.m:
                self.m[k] = torch.zeros_like(g)
                self.v[k] = torch.zeros_like(g)
            p = params[k]
            p.mul_(1.0 - self.lr * self.wd)
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(1 - b2 ** self.t)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / (1 - b1 ** self.t))
        return total


def train_step(fus: Params, clf: Params, batch: Dict[str, Tensor], opt: AdamWState, dropout: float = 0.1,
               masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """One iteration of ForensicTrainer._epoch_loop (forensic_trainer.py:285-298), in place on fus/clf."""
    out, gf, gc = loss_and_grads(fus, clf, batch, dropout, masks)
    grads = {("fusion." + k): g for k, g in gf.items()}
    grads.update({("clf." + k): g for k, g in gc.items()})
    params = {("fusion." + k): fus[k] for k in gf}
    params.update({("clf." + k): clf[k] for k in gc})
    out["grad_norm"] = opt.step(params, grads)
    return out


class TorchStep:
    """The reference's OWN step machinery around the restated forward — leaf parameters, ``F.cross_entropy``,
    ``loss.backward()``, ``nn.utils.clip_grad_norm_(params, 5.0)`` and ``torch.optim.AdamW(lr=2e-4, weight_decay=1e-4)``
    exactly as src/training/forensic_trainer.py:173-177,286-298 drives them — on any device. Used as the timed CPU
    baseline of bench.py (no per-step parameter clones, torch's own multi-tensor optimizer: what the reference runs)
    and, on ``cuda``, as the same-box "stock PyTorch eager" bar. ``AdamWState`` above stays the
    spelled-out arithmetic that the golden trajectories pin; tests/test_oracle_golden.py checks the two agree."""

    def __init__(self, fus: Params, clf: Params, device: str = "cpu", lr: float = 2e-4, weight_decay: float = 1e-4,
                 grad_clip: float = 5.0):
        fk, ck = trainable_keys()
        self.fus = {k: v.detach().to(device).clone().requires_grad_(k in fk) for k, v in fus.items()}
        self.clf = {k: v.detach().to(device).clone().requires_grad_(k in ck) for k, v in clf.items()}
        self.params = [self.fus[k] for k in fk] + [self.clf[k] for k in ck]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.clip = grad_clip

    def step(self, batch: Dict[str, Tensor], dropout: float = 0.1, masks: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
        out = model_forward(self.fus, self.clf, batch, dropout, masks)
        self.opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        if self.clip and self.clip > 0:
            out["grad_norm"] = torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.clip)
        self.opt.step()
        return out


def rel_err(a: Tensor, b: Tensor) -> float:
    """Norm-wise relative error ||a-b|| / ||b|| used by every parity test."""
    a, b = a.double().flatten(), b.double().flatten()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)
