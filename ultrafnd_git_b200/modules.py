"""Drop-in replacements for the reference's fusion modules, backed by libfnd_b200.so.

Mirrors (names, constructor arguments, forward dict in/out, state_dict keys, attributes) of
  * ``CrossModalTransformer`` / ``ForensicCoAttention``  — src/models/fusion/cross_modal_transformer.py:17-210
  * ``DeepTruthClassifier`` / ``NODEEnsemble`` / ``_ObliviousTree`` — src/models/fusion/deep_truth_classifier.py:28-184

The nn.Modules here are parameter CONTAINERS: every parameter is a view into the engine's flat fp32 arena, and
``forward`` hands raw device pointers to the C ABI (``fnd_fusion_forward`` / ``fnd_classifier_forward``) through a
``torch.autograd.Function`` whose backward calls ``fnd_*_backward``. There is no eager/CPU compute path: calling
``forward`` without a CUDA device raises.

Differences from the reference, all deliberate and documented in DESIGN.md:
  * the modules live on the current CUDA device (the reference pins itself to mps|cpu and cannot run on CUDA);
  * ``probs`` and the fusion head's ``logits`` are returned but not differentiable (the reference training step
    never differentiates them: forensic_trainer.py:261-267,287);
  * gradients w.r.t. the input feature tensors are not produced (they are data in the reference trainer).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as E
from ._lib import check
from .config_utils import ConfigManager


class _ArenaLinear(nn.Module):
    """Parameter holder shaped like ``nn.Linear`` (weight [out,in], bias [out]). Calling it directly uses torch
    (convenience for inspection only — the hot path never does)."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        lin = nn.Linear(in_features, out_features)          # consumes the RNG exactly like the reference
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features} (arena-backed)"


def _xavier_(m: _ArenaLinear) -> None:
    nn.init.xavier_uniform_(m.weight)
    nn.init.zeros_(m.bias)


class ForensicCoAttention(nn.Module):
    """Parameter container for one evidence-gated co-attention block (cross_modal_transformer.py:27-37)."""

    def __init__(self, hidden_dim: int, evidence_dim: int = 3):
        super().__init__()
        self.h = hidden_dim
        self.q = _ArenaLinear(hidden_dim, hidden_dim)
        self.k = _ArenaLinear(hidden_dim, hidden_dim)
        self.v = _ArenaLinear(hidden_dim, hidden_dim)
        self.evidence_proj = nn.Sequential(_ArenaLinear(evidence_dim, hidden_dim), nn.GELU(), _ArenaLinear(hidden_dim, 1))


class _SemanticStub(nn.Module):
    """Keeps the never-used ``semantic.*`` projection weights in the state_dict: the reference
    constructs SemanticForgeryAnalyzer inside the fusion module, so its two Linear(512,512) appear in checkpoints."""

    def __init__(self):
        super().__init__()
        self.text_proj = nn.Sequential(_ArenaLinear(512, 512))
        self.vision_proj = nn.Sequential(_ArenaLinear(512, 512))


class _EngineModule(nn.Module):
    """Shared machinery: parameters are re-pointed into an Engine arena; ``.to()`` / ``.cuda()`` keep them there."""
    _prefix = ""

    def _attach(self, eng: E.Engine) -> None:
        object.__setattr__(self, "_engine", eng)
        eng.attached.append(weakref.ref(self))
        with torch.no_grad():
            for name, p in self.named_parameters():
                view = eng.view(self._prefix + name)
                if p.data.data_ptr() != view.data_ptr():
                    view.copy_(p.data.to(view.device, torch.float32))
                    p.data = view
        eng._shadow_version = None

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, dtype=torch.float32, device=self._engine.device))
        if probe.dtype != torch.float32:
            raise NotImplementedError("ultrafnd_git_b200 keeps fp32 master parameters; choose the compute precision "
                                      "with FND_PRECISION / precision= instead of .half()/.double()")
        if probe.device != self._engine.device:
            self._engine.to(probe.device)
            with torch.no_grad():
                for name, p in self.named_parameters():
                    p.data = self._engine.view(self._prefix + name)
        return self

    def _param_version(self) -> int:
        return sum(p._version for p in self.parameters())

    @property
    def device(self) -> torch.device:
        return self._engine.device

    @device.setter
    def device(self, value) -> None:     # the reference exposes a writable attribute (test harnesses patch it)
        self.to(torch.device(value))

    @property
    def precision(self) -> str:
        return "fp32" if self._engine.mode == E.MODE_FP32X3 else "bf16"

    def set_precision(self, precision: str) -> "._EngineModule":
        self._engine.set_mode(E.MODE_FP32X3 if precision in ("fp32", "fp32x3") else E.MODE_BF16)
        return self


# ======================================================================================================
# CrossModalTransformer
# ======================================================================================================
class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, t, a, v, u, g, *params):
        eng: E.Engine = module._engine
        B = t.shape[0]
        plan = eng.plan(B)
        eng.ensure_shadows(module._shadow_key())
        feats = {"text_features": t, "audio_features": a, "visual_features": v, "temporal_features": u, "gnn_feat": g}
        inp, keep = E.make_inputs(feats, use_gnn=eng.dims.use_gnn)
        training = bool(module.training and (eng.dims.fusion_dropout > 0))
        check(eng.lib.fnd_fusion_forward(plan.handle, ctypes.byref(inp), int(training), eng.stream_ptr()), "fnd_fusion_forward")
        plan.fusion_id += 1
        H = eng.dims.hidden
        fused = plan.buffer("fused", torch.float32, (B, H)).clone()
        logits = plan.buffer("fusion_logits", torch.float32, (B, 2)).clone()
        rs = plan.buffer("rowstat", torch.float32, (B, 16))
        sc, emo, delay = rs[:, 0].clone(), rs[:, 1].clone(), rs[:, 2].clone()
        ctx.module, ctx.plan, ctx.fid, ctx.nparams = module, plan, plan.fusion_id, len(params)
        ctx.mark_non_differentiable(logits, sc, emo, delay)
        return fused, logits, sc, emo, delay

    @staticmethod
    def backward(ctx, dfused, *_unused):
        module, plan = ctx.module, ctx.plan
        eng: E.Engine = module._engine
        if plan.fusion_id != ctx.fid:
            raise RuntimeError("CrossModalTransformer backward: the activations saved in this batch size's plan were "
                               "overwritten by a later forward; run backward before the next forward of the same batch size")
        dfused = dfused.to(torch.float32).contiguous()
        check(eng.lib.fnd_fusion_backward(plan.handle, dfused.data_ptr(), None, eng.stream_ptr()), "fnd_fusion_backward")
        flat = eng.grads.clone()       # autograd may keep/accumulate the returned tensors; the arena is overwritten next step
        grads = tuple(eng.grad_view(module._prefix + n, flat) for n in module._param_names)
        # Gradients w.r.t. the modality vectors (only when they come from a trainable SequenceFrontEnd): dX_m = dP_m W_m.
        # dP (bf16, [B, 5H], the operand the projection wgrad reads) is still in the plan's workspace; the product runs
        # on the library's persistent tcgen05 GEMM. The evidence scalars are computed under no_grad in the reference
        # (cross_modal_transformer.py:153-164), so the projections are the only path back to the inputs.
        dins = [None, None, None, None]
        if any(ctx.needs_input_grad[1:5]):
            from . import seq_ops as S
            B, H = dfused.shape[0], eng.dims.hidden
            planes = [plan.buffer("dP_hi", torch.bfloat16, (B, 5 * H))]
            try:
                planes.append(plan.buffer("dP_lo", torch.bfloat16, (B, 5 * H)))        # fp32 mode: dP = hi + lo
            except KeyError:
                pass
            projs = (module.text_proj, module.audio_proj, module.visual_proj, module.temporal_proj)
            for m, proj in enumerate(projs):
                if not ctx.needs_input_grad[1 + m]:
                    continue
                wt = proj.weight.detach().to(torch.bfloat16).t().contiguous()          # [D_in, H]
                dx = None
                for dP in planes:
                    part = torch.empty(B, wt.shape[0], dtype=torch.float32, device=dfused.device)
                    S.linear(dP[:, m * H:(m + 1) * H], wt, None, out_f32=part, want_bf16=False)
                    dx = part if dx is None else dx + part
                dins[m] = dx
        return (None, dins[0], dins[1], dins[2], dins[3], None) + grads


class CrossModalTransformer(_EngineModule):
    """B200 drop-in for the reference fusion module (cross_modal_transformer.py:63-210).

    forward(feats) with ``text_features (B,768)``, ``audio_features (B,128)``, ``visual_features (B,512)``,
    ``temporal_features (B,256)`` and ``gnn_feat (B,gnn_dim)`` returns
    ``{"fused": (B,H), "logits": (B,2), "forensic": {emotion_intensity, semantic_conflict, temporal_delay}}``.
    """
    _prefix = "fusion."

    def __init__(self, config_path: str = "configs/model_configs/fusion.yaml", precision: Optional[str] = None,
                 device: Optional[torch.device] = None):
        super().__init__()
        cfg = ConfigManager().load_config(config_path)
        self.hidden = int(cfg.get("hidden_dim", 512))
        self.dropout = float(cfg.get("dropout", 0.3))
        self.use_gnn = bool(cfg.get("use_gnn", True))
        self.gnn_dim = int(cfg.get("gnn_dim", 128))
        self.dtype = torch.float32
        H = self.hidden
        # registration order == the reference's, so state_dict order and the init RNG stream match
        self.text_proj = _ArenaLinear(768, H)
        self.audio_proj = _ArenaLinear(128, H)
        self.visual_proj = _ArenaLinear(512, H)
        self.temporal_proj = _ArenaLinear(256, H)
        if self.use_gnn:
            self.gnn_proj = _ArenaLinear(self.gnn_dim, H)
        self.semantic = _SemanticStub()
        self.attn_tv = ForensicCoAttention(H, evidence_dim=3)
        self.attn_ta = ForensicCoAttention(H, evidence_dim=3)
        self.attn_vu = ForensicCoAttention(H, evidence_dim=3)
        self.include_pairs = True
        self.fused_dim = (4 + 8 + 3 + (1 if self.use_gnn else 0)) * H
        self.fuse_mlp = nn.Sequential(_ArenaLinear(self.fused_dim, 2 * H), nn.GELU(), nn.Dropout(self.dropout),
                                      _ArenaLinear(2 * H, H), nn.GELU(), nn.Dropout(self.dropout))
        self.classifier = _ArenaLinear(H, 2)
        dims = E.Dims(hidden=H, d_gnn=self.gnn_dim, use_gnn=self.use_gnn, fusion_dropout=self.dropout)
        mode = None if precision is None else (E.MODE_FP32X3 if precision in ("fp32", "fp32x3") else E.MODE_BF16)
        self._param_names = [n for n, _ in self.named_parameters()]
        self._attach(E.Engine(dims, device=device, mode=mode))

    def _shadow_key(self) -> int:
        return self._engine.param_version()

    def _sync_dropout(self) -> None:
        p = float(self.fuse_mlp[2].p)
        if p != self._engine.dims.fusion_dropout:
            self._engine.set_dropout(fusion_p=p)

    def attach_sequence_frontend(self, frontend) -> None:
        """Tier B hook: with a ``SequenceFrontEnd`` attached, ``forward`` accepts 3-D feature
        tensors (B, L, D_in) (+ optional ``<key>_mask``) and pools them to the (B, D) vectors first; 2-D inputs behave
        exactly as before. Kept out of ``state_dict`` (the reference's checkpoint schema has no such entries)."""
        object.__setattr__(self, "_seq_frontend", frontend)

    def forward(self, feats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        eng = self._engine
        eng.require_cuda()
        self._sync_dropout()
        if feats["text_features"].dim() == 3:
            fe = getattr(self, "_seq_frontend", None)
            if fe is None:
                raise RuntimeError("3-D (sequence) features need a SequenceFrontEnd: call attach_sequence_frontend() first")
            feats = fe.forward_features(feats)

        def prep(x: torch.Tensor) -> torch.Tensor:
            # the four modality vectors may carry a graph (a trainable SequenceFrontEnd produced them): _FusionFn.backward
            # returns dL/dx for them; gnn_feat never does
            return x.to(eng.device, dtype=torch.float32)

        t, a = prep(feats["text_features"]), prep(feats["audio_features"])
        v, u = prep(feats["visual_features"]), prep(feats["temporal_features"])
        g = None
        if self.use_gnn:
            if feats.get("gnn_feat") is None:
                # mirrors the reference: fuse_mlp.0 was sized for 16 slots at construction, so a missing gnn_feat
                # is a shape error there too (cross_modal_transformer.py:184,197; SURVEY.md §7 hard parts)
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({t.shape[0]}x{15 * self.hidden} and "
                                   f"{self.fused_dim}x{2 * self.hidden}): gnn_feat is required when use_gnn=True")
            if feats["gnn_feat"].requires_grad:
                raise NotImplementedError("gradients w.r.t. gnn_feat are not produced by this drop-in")
            g = prep(feats["gnn_feat"])
        params = tuple(self.parameters())
        fused, logits, sc, emo, delay = _FusionFn.apply(self, t, a, v, u, g, *params)
        return {"fused": fused, "logits": logits,
                "forensic": {"emotion_intensity": emo, "semantic_conflict": sc, "temporal_delay": delay}}


# ======================================================================================================
# DeepTruthClassifier
# ======================================================================================================
class _ObliviousTree(nn.Module):
    """Parameter container for one soft oblivious tree (deep_truth_classifier.py:36-52)."""

    def __init__(self, in_dim: int, num_classes: int = 2, depth: int = 4, tau: float = 10.0, dropout: float = 0.3):
        super().__init__()
        self.in_dim, self.depth, self.num_classes = in_dim, depth, num_classes
        self.tau = nn.Parameter(torch.tensor(float(tau)), requires_grad=False)
        self.gates = nn.ParameterList([nn.Parameter(torch.zeros(in_dim)) for _ in range(depth)])
        self.thresh = nn.ParameterList([nn.Parameter(torch.zeros(1)) for _ in range(depth)])
        self.num_leaves = 1 << depth
        self.leaf_logits = nn.Parameter(torch.zeros(self.num_leaves, num_classes))
        self.dropout = nn.Dropout(dropout)


class NODEEnsemble(nn.Module):
    def __init__(self, in_dim: int, num_classes: int = 2, num_trees: int = 6, depth: int = 4, tau: float = 10.0,
                 dropout: float = 0.3):
        super().__init__()
        self.trees = nn.ModuleList([_ObliviousTree(in_dim, num_classes, depth=depth, tau=tau, dropout=dropout)
                                    for _ in range(num_trees)])


class _ClassifierFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, fused, aux, *params):
        eng: E.Engine = module._engine
        B = fused.shape[0]
        plan = eng.plan(B)
        eng.ensure_shadows(module._shadow_key())
        fused_c = fused.detach().to(torch.float32).contiguous()
        aux_c = aux.detach().to(torch.float32).contiguous() if aux is not None else None
        training = bool(module.training)
        check(eng.lib.fnd_classifier_forward(plan.handle, fused_c.data_ptr(), aux_c.data_ptr() if aux_c is not None else None,
                                             aux_c.shape[1] if aux_c is not None else 0, int(training), eng.stream_ptr()),
              "fnd_classifier_forward")
        plan.clf_id += 1
        logits = plan.buffer("logits", torch.float32, (B, 2)).clone()
        probs = plan.buffer("probs", torch.float32, (B, 2)).clone()
        ctx.module, ctx.plan, ctx.fid = module, plan, plan.clf_id
        ctx.need_daux = bool(aux is not None and aux.requires_grad)
        ctx.mark_non_differentiable(probs)
        return logits, probs

    @staticmethod
    def backward(ctx, dlogits, _dprobs):
        module, plan = ctx.module, ctx.plan
        eng: E.Engine = module._engine
        if plan.clf_id != ctx.fid:
            raise RuntimeError("DeepTruthClassifier backward: the activations saved in this batch size's plan were "
                               "overwritten by a later forward; run backward before the next forward of the same batch size")
        dlogits = dlogits.to(torch.float32).contiguous()
        check(eng.lib.fnd_classifier_backward(plan.handle, dlogits.data_ptr(), eng.stream_ptr()), "fnd_classifier_backward")
        B = dlogits.shape[0]
        H = eng.dims.hidden
        dfused = plan.buffer("dfused", torch.float32, (B, H)).clone()
        daux = None
        if ctx.need_daux:
            # d loss / d aux = dz_pre0 . W_pre0[:, H:H+aux_dim] (deep_truth_classifier.py:142-146: aux is concatenated behind
            # fused). dz_pre0 is the backward's own intermediate (bf16 hi [+ lo in fp32 mode]); two output columns, so a
            # broadcast-multiply-reduce rather than a GEMM.
            dz = plan.buffer("dz_p0_hi", torch.bfloat16, (B, H)).float()
            if eng.mode == E.MODE_FP32X3:
                dz = dz + plan.buffer("dz_p0_lo", torch.bfloat16, (B, H)).float()
            w_aux = module.pre[0].weight.detach()[:, H:H + eng.dims.aux_dim]
            daux = (dz[:, :, None] * w_aux[None, :, :]).sum(dim=1)
        flat = eng.grads.clone()
        grads = tuple(eng.grad_view(module._prefix + n, flat) for n in module._param_names)
        return (None, dfused, daux) + grads


class DeepTruthClassifier(_EngineModule):
    """B200 drop-in for the reference classifier (deep_truth_classifier.py:97-184).

    forward(fused (B,F), aux (B,2) or None) -> {"logits": (B,2), "probs": (B,2), "temperature": scalar}.
    """
    _prefix = "clf."

    def __init__(self, config_path: str = "configs/model_configs/classifier.yaml", precision: Optional[str] = None,
                 device: Optional[torch.device] = None):
        super().__init__()
        cfg = ConfigManager().load_config(config_path)
        self.hidden = int(cfg.get("hidden_dim", 512))
        self.dropout = float(cfg.get("dropout", 0.3))
        self.num_classes = int(cfg.get("num_classes", 2))
        self.use_aux = bool(cfg.get("use_aux", True))
        self.aux_dim = int(cfg.get("aux_dim", 2))
        self.node_trees = int(cfg.get("node_trees", 6))
        self.node_depth = int(cfg.get("node_depth", 4))
        self.node_tau = float(cfg.get("node_tau", 10.0))
        self.temperature = nn.Parameter(torch.tensor(float(cfg.get("temperature", 1.0))), requires_grad=True)
        in_dim = int(cfg.get("input_dim", self.hidden))
        if self.num_classes != 2:
            raise NotImplementedError("num_classes != 2 is not supported by the sm_100a head kernel")
        if in_dim != self.hidden:
            raise NotImplementedError("classifier input_dim must equal hidden_dim (the reference default, classifier.yaml:2-3)")
        eff_in = in_dim + (self.aux_dim if self.use_aux else 0)
        self.pre = nn.Sequential(_ArenaLinear(eff_in, self.hidden), nn.GELU(), nn.Dropout(self.dropout),
                                 _ArenaLinear(self.hidden, self.hidden), nn.GELU(), nn.Dropout(self.dropout))
        for m in self.pre:
            if isinstance(m, _ArenaLinear):
                _xavier_(m)
        self.node = NODEEnsemble(in_dim=self.hidden, num_classes=self.num_classes, num_trees=self.node_trees,
                                 depth=self.node_depth, tau=self.node_tau, dropout=0.3)
        self.bypass = _ArenaLinear(self.hidden, self.num_classes)
        _xavier_(self.bypass)
        dims = E.Dims(hidden=self.hidden, aux_dim=self.aux_dim if self.use_aux else 0, trees=self.node_trees,
                      depth=self.node_depth, clf_dropout=self.dropout, tree_dropout=0.3, node_tau=self.node_tau)
        mode = None if precision is None else (E.MODE_FP32X3 if precision in ("fp32", "fp32x3") else E.MODE_BF16)
        self._param_names = [n for n, _ in self.named_parameters()]
        self._attach(E.Engine(dims, device=device, mode=mode))

    def _shadow_key(self) -> int:
        return self._engine.param_version()

    def _sync_dropout(self) -> None:
        p = float(self.pre[2].p)
        tp = float(self.node.trees[0].dropout.p)
        d = self._engine.dims
        if p != d.clf_dropout or tp != d.tree_dropout:
            self._engine.set_dropout(clf_p=p, tree_p=tp)

    def forward(self, fused: torch.Tensor, aux: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        eng = self._engine
        eng.require_cuda()
        self._sync_dropout()
        fused = fused.to(eng.device, dtype=torch.float32)
        if self.use_aux:
            if aux is None:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({fused.shape[0]}x{self.hidden} and "
                                   f"{self.hidden + self.aux_dim}x{self.hidden}): aux is required when use_aux=True")
            aux = aux.to(eng.device, dtype=torch.float32)
        else:
            aux = None
        params = tuple(self.parameters())
        logits, probs = _ClassifierFn.apply(self, fused, aux, *params)
        t = torch.clamp(self.temperature.detach(), min=0.5, max=5.0)
        return {"logits": logits, "probs": probs, "temperature": t}

    def feature_importance(self, fused: torch.Tensor, aux: Optional[torch.Tensor] = None, class_idx: int = 1,
                           aggregate: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Gradient x Input importance (deep_truth_classifier.py:189-211): |d logits[:, class_idx].sum() / d x * x| with
        x = [fused | aux]. Returns (per-input importance (B, F[+A]), its batch mean (F[+A],) when ``aggregate``). The input
        gradients come from the library's own backward (``dfused`` crosses the C ABI; d/d aux from its dz_pre0)."""
        dev = self._engine.device
        fused = fused.detach().to(dev, dtype=torch.float32).requires_grad_(True)
        aux = aux.detach().to(dev, dtype=torch.float32).requires_grad_(True) if (aux is not None and self.use_aux) else None
        logits = self.forward(fused, aux)["logits"]
        logits[:, class_idx].sum().backward()
        x = torch.cat([fused, aux], dim=-1) if aux is not None else fused
        grad = torch.cat([fused.grad, aux.grad], dim=-1) if aux is not None else fused.grad
        imp = (grad.detach() * x.detach()).abs()
        return (imp, imp.mean(dim=0)) if aggregate else (imp, None)

    @torch.no_grad()
    def predict_proba(self, fused: torch.Tensor, aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.forward(fused, aux)["probs"]

    @torch.no_grad()
    def predict(self, fused: torch.Tensor, aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.predict_proba(fused, aux).argmax(dim=-1)


# ======================================================================================================
# Pairing: one arena for both modules (what the fused trainer step needs)
# ======================================================================================================
def pair_modules(fusion: CrossModalTransformer, clf: DeepTruthClassifier, precision: Optional[str] = None) -> E.Engine:
    """Move both modules onto ONE engine (one arena, one set of plans) so ``fnd_train_step`` / ``fnd_eval_step`` can
    run the whole ForensicTrainer._forward_batch + loss + backward + optimizer sequence without leaving the library."""
    if fusion._engine is clf._engine:
        return fusion._engine
    fd, cd = fusion._engine.dims, clf._engine.dims
    if fd.hidden != cd.hidden:
        raise NotImplementedError("fusion hidden_dim must equal classifier hidden_dim / input_dim")
    dims = E.Dims(hidden=fd.hidden, d_gnn=fd.d_gnn, use_gnn=fd.use_gnn, fusion_dropout=fd.fusion_dropout,
                  aux_dim=cd.aux_dim, trees=cd.trees, depth=cd.depth, clf_dropout=cd.clf_dropout,
                  tree_dropout=cd.tree_dropout, node_tau=cd.node_tau)
    mode = fusion._engine.mode if precision is None else (E.MODE_FP32X3 if precision in ("fp32", "fp32x3") else E.MODE_BF16)
    eng = E.Engine(dims, device=fusion._engine.device, mode=mode)
    fusion._attach(eng)
    clf._attach(eng)
    return eng
