"""Fused trainer step: ForensicTrainer._forward_batch + F.cross_entropy + backward + clip_grad_norm_ + AdamW
(src/training/forensic_trainer.py:238-298) as ONE stream-ordered launch sequence inside libfnd_b200.so, optionally
replayed from a CUDA graph.

``FusedStep`` owns static device input buffers (so a captured graph can be replayed while a copy stream refills
them) or, alternatively, gathers each batch from a device-resident feature cache by row index.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch

from . import engine as E
from ._lib import FndInputs, check
from .modules import CrossModalTransformer, DeepTruthClassifier, pair_modules

FEATURE_KEYS = ("text_features", "audio_features", "visual_features", "temporal_features", "gnn_feat")


class DeviceCache:
    """The whole feature cache as one fp32 matrix on the device: columns [text | audio | visual | temporal | gnn | aux]
    with a 16-byte-aligned row pitch, plus int64 labels. Replaces CachedTensorDataset + default collate +
    ``gnn_Z[global_idx]`` + ``.to(device)`` per step (forensic_trainer.py:60-83,238-263)."""

    def __init__(self, feats: Dict[str, torch.Tensor], aux: torch.Tensor, labels: torch.Tensor, device: torch.device):
        widths = [feats[k].shape[1] for k in FEATURE_KEYS]
        self.offsets = [0]
        for w in widths:
            self.offsets.append(self.offsets[-1] + w)
        self.aux_off = self.offsets[-1]
        total = self.aux_off + 2
        self.pitch = (total + 3) // 4 * 4
        n = labels.shape[0]
        self.n = n
        m = torch.zeros(n, self.pitch, dtype=torch.float32)
        for k, o in zip(FEATURE_KEYS, self.offsets):
            m[:, o:o + feats[k].shape[1]] = feats[k].to(torch.float32)
        m[:, self.aux_off:self.aux_off + 2] = aux.to(torch.float32)
        self.matrix = m.to(device)
        self.labels = labels.to(torch.int64).to(device)
        self.widths = widths

    def inputs(self, gather: torch.Tensor) -> FndInputs:
        inp = FndInputs()
        base = self.matrix.data_ptr()
        for i in range(5):
            inp.x[i] = base + 4 * self.offsets[i]
            inp.pitch[i] = self.pitch
        inp.aux = base + 4 * self.aux_off
        inp.aux_pitch = self.pitch
        inp.labels = self.labels.data_ptr()
        inp.gather = gather.data_ptr()
        return inp


class FusedStep:
    """Runs fnd_train_step / fnd_eval_step for a (fusion, classifier) pair at a fixed batch size."""

    def __init__(self, fusion: CrossModalTransformer, clf: DeepTruthClassifier, batch: int,
                 precision: Optional[str] = None, use_graph: bool = True, dp_group=None):
        """dp_group: a torch.distributed process group (one rank per GPU of one NVSwitch domain). When given, rank 0's
        parameters are broadcast, the arena moves into peer-mapped memory and ``dp_optimizer_step`` becomes available."""
        self.fusion, self.clf = fusion, clf
        self.engine = pair_modules(fusion, clf, precision)
        self.engine.require_cuda()
        self.engine.enable_optimizer()
        if dp_group is not None and self.engine.symm is None:
            import torch.distributed as dist
            dist.broadcast(self.engine.params, src=dist.get_global_rank(dp_group, 0), group=dp_group)
            self.engine.enable_symmetric(dp_group)
        self.batch = batch
        self.plan = self.engine.plan(batch)
        self.use_graph = use_graph
        # FND_DP_OVERLAP=1: push the fuse_mlp.0/.3 gradients from a side stream under the rest of the backward pass.
        # Off by default: on this pool's NVSwitch boxes the one-CTA-per-SM push that fits beside the backward's GEMM CTAs
        # reaches ~160 GB/s and finishes after the backward, while the full-grid push that follows it takes 40 us.
        self.dp_overlap = os.environ.get("FND_DP_OVERLAP", "0") == "1"
        # FND_DP_DEFER=1: defer the update + all-gather of the fuse_mlp slice to the next step's side stream (under its
        # first four kernels; flushed automatically before any other entry point, lr change or state_dict read). Off by
        # default: measured on 8 GPUs the side-stream kernel slows the latency-bound forward chain as much as it saves
        # (321 vs 316 us/step), like the early push.
        self.dp_defer = os.environ.get("FND_DP_DEFER", "0") == "1"
        # FND_DP_FUSED=1 (default; bf16 wire format only): the weight-gradient launch's epilogue stores every tile straight
        # into its owner's staging slot over NVLink (no separate push kernel, the fp32 gradient never touches HBM).
        self.dp_fused = os.environ.get("FND_DP_FUSED", "1") == "1"
        # FND_WG_EARLY=1: single-GPU step with the fuse_mlp weight gradients launched early on a side stream
        self.wg_early = os.environ.get("FND_WG_EARLY", "0") == "1"
        self._side_stream: Optional[torch.cuda.Stream] = None
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        dev = self.engine.device
        d = self.engine.dims
        widths = [d.d_text, d.d_audio, d.d_visual, d.d_temporal, d.d_gnn]
        # static staging buffers: [B, sum(widths)+2] fp32 (+ labels, gather indices)
        self.in_off = [0]
        for w in widths:
            self.in_off.append(self.in_off[-1] + w)
        self.aux_off = self.in_off[-1]
        self.in_pitch = (self.aux_off + 2 + 3) // 4 * 4
        self.static_in = torch.zeros(batch, self.in_pitch, dtype=torch.float32, device=dev)
        self.static_labels = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.static_gather = torch.zeros(batch, dtype=torch.int64, device=dev)
        # second index buffer: the row indices of step i+1 are staged (side stream) while the graph of step i reads the other
        self.alt_gather = torch.zeros(batch, dtype=torch.int64, device=dev)
        self._inp_static = self._make_static_inputs()
        # A second input set for double-buffered host feeding: the H2D copy of batch i+1 lands in one set while the
        # graph of batch i reads the other (no device-to-device staging copy on the critical path).
        self.alt_in = torch.zeros_like(self.static_in)
        self.alt_labels = torch.zeros_like(self.static_labels)
        self._inp_alt = self._make_static_inputs(self.alt_in, self.alt_labels)
        self._inp_cache: Optional[FndInputs] = None
        self._inp_cache_alt: Optional[FndInputs] = None
        self._cache: Optional[DeviceCache] = None
        self.engine.refresh_shadows(self.engine.param_version())

    # ------------------------------------------------------------------ inputs
    def _make_static_inputs(self, buf: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None) -> FndInputs:
        buf = self.static_in if buf is None else buf
        labels = self.static_labels if labels is None else labels
        inp = FndInputs()
        base = buf.data_ptr()
        for i in range(5):
            inp.x[i] = base + 4 * self.in_off[i]
            inp.pitch[i] = self.in_pitch
        inp.aux = base + 4 * self.aux_off
        inp.aux_pitch = self.in_pitch
        inp.labels = labels.data_ptr()
        inp.gather = None
        return inp

    def gather_set(self, which: int) -> torch.Tensor:
        """Row-index buffer (int64 [B]) read by the cache-fed graphs of input set 0 / 1."""
        return self.static_gather if which == 0 else self.alt_gather

    def input_set(self, which: int):
        """(inputs, labels) device buffers of input set 0 / 1 (layout of host_staging())."""
        return (self.static_in, self.static_labels) if which == 0 else (self.alt_in, self.alt_labels)

    def host_staging(self) -> Dict[str, torch.Tensor]:
        """Pinned host buffers with the static device buffers' layout (for the end-to-end path)."""
        return {"inputs": torch.zeros(self.batch, self.in_pitch, dtype=torch.float32).pin_memory(),
                "labels": torch.zeros(self.batch, dtype=torch.int64).pin_memory()}

    def pack_host(self, batch: Dict[str, torch.Tensor], staging: Dict[str, torch.Tensor]) -> None:
        m = staging["inputs"]
        for k, o in zip(FEATURE_KEYS, self.in_off):
            m[:, o:o + batch[k].shape[1]] = batch[k]
        m[:, self.aux_off:self.aux_off + 2] = batch["aux"]
        staging["labels"].copy_(batch["label"])

    def load_batch(self, batch: Dict[str, torch.Tensor]) -> None:
        """Copy one batch of (device or host) tensors into the static input buffers."""
        for k, o in zip(FEATURE_KEYS, self.in_off):
            self.static_in[:, o:o + batch[k].shape[1]].copy_(batch[k], non_blocking=True)
        self.static_in[:, self.aux_off:self.aux_off + 2].copy_(batch["aux"], non_blocking=True)
        if "label" in batch:
            self.static_labels.copy_(batch["label"], non_blocking=True)

    def upload(self, staging: Dict[str, torch.Tensor]) -> None:
        self.static_in.copy_(staging["inputs"], non_blocking=True)
        self.static_labels.copy_(staging["labels"], non_blocking=True)

    def attach_cache(self, cache: DeviceCache) -> None:
        self._cache = cache
        self._inp_cache = cache.inputs(self.static_gather)
        self._inp_cache_alt = cache.inputs(self.alt_gather)
        self._graphs.clear()

    # ------------------------------------------------------------------ launches
    def _run(self, entry: str, inp: FndInputs) -> None:
        lib, h = self.engine.lib, self.plan.handle
        if entry == "train_step_dp":      # forward + backward + the sharded peer-memory optimizer step
            flags = (1 if self.dp_overlap else 0) | (2 if self.dp_defer else 0)
            side = None
            fused = 4 if (self.dp_fused and not self.dp_overlap) else 0
            if flags:
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(self.engine.device)
                side = self._side_stream.cuda_stream
            check(lib.fnd_train_step_dp(h, ctypes.byref(inp), self.engine.stream_ptr(), side, flags | fused), "fnd_train_step_dp")
            return
        if entry == "train_step" and self.wg_early:
            # fuse_mlp weight gradients on a side stream under the rest of the backward chain (fork / join inside the call)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(self.engine.device)
            check(lib.fnd_train_step_overlap(h, ctypes.byref(inp), self.engine.stream_ptr(), self._side_stream.cuda_stream),
                  "fnd_train_step_overlap")
            return
        fn = {"train_step": lib.fnd_train_step, "train_fwd_bwd": lib.fnd_train_fwd_bwd, "eval_step": lib.fnd_eval_step}[entry]
        check(fn(h, ctypes.byref(inp), self.engine.stream_ptr()), "fnd_" + entry)

    def _launch(self, entry: str, from_cache: bool, input_set: int = 0) -> None:
        if from_cache:
            inp = self._inp_cache if input_set == 0 else self._inp_cache_alt
        else:
            inp = self._inp_static if input_set == 0 else self._inp_alt
        if inp is None:
            raise RuntimeError("attach_cache() first")
        if entry != "train_step_dp":
            self.engine.dp_flush()        # a deferred data-parallel update must land before anything else reads the model
        if not self.use_graph:
            self._run(entry, inp)
            return
        key = entry + (("/cache" if input_set == 0 else "/cache1") if from_cache else ("/static" if input_set == 0 else "/alt"))
        g = self._graphs.get(key)
        if g is None:
            # warm-up on a side stream (first launches set function attributes), then capture
            s = torch.cuda.Stream(self.engine.device)
            s.wait_stream(torch.cuda.current_stream(self.engine.device))
            with torch.cuda.stream(s):
                self._run("eval_step", inp)
            torch.cuda.current_stream(self.engine.device).wait_stream(s)
            torch.cuda.synchronize(self.engine.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(entry, inp)
            self._graphs[key] = g
        g.replay()

    def train_step(self, from_cache: bool = False, input_set: int = 0) -> None:
        """One optimizer step on the batch currently in the static buffers of `input_set` (or gathered from the cache)."""
        self._launch("train_step", from_cache, input_set)
        self.plan.bump_all()

    def train_step_dp(self, from_cache: bool = False, input_set: int = 0) -> None:
        """One data-parallel optimizer step (every rank calls it with its own batch slice): forward + backward +
        peer-memory reduce-scatter / sharded AdamW / shadow all-gather, ONE CUDA graph per rank."""
        if not getattr(self, "_dp_bound", False):
            self.engine.dp_bind(self.plan)
            self._dp_bound = True
        self.engine._dp_plan = self.plan
        self._launch("train_step_dp", from_cache, input_set)
        self.engine._dp_pending = self.dp_defer
        self.plan.bump_all()

    def train_fwd_bwd(self, from_cache: bool = False, input_set: int = 0) -> None:
        """Forward + loss + backward only (gradients left in the arena for an all-reduce)."""
        self._launch("train_fwd_bwd", from_cache, input_set)
        self.plan.bump_all()

    def dp_optimizer_step(self) -> None:
        """Sharded clip + AdamW over NVLink peer memory (Engine.enable_symmetric must have been called): the gradients
        left by train_fwd_bwd on every rank are reduce-scattered, the slice is updated and the bf16 shadows are written
        to all ranks. Replayed from a CUDA graph together with nothing else (the kernels spin on peer flags)."""
        if not getattr(self, "_dp_bound", False):
            self.engine.dp_bind(self.plan)
            self._dp_bound = True
        if not self.use_graph:
            check(self.engine.lib.fnd_dp_optimizer_step(self.plan.handle, self.engine.stream_ptr()), "fnd_dp_optimizer_step")
            return
        g = self._graphs.get("dp_opt")
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                check(self.engine.lib.fnd_dp_optimizer_step(self.plan.handle, self.engine.stream_ptr()), "fnd_dp_optimizer_step")
            self._graphs["dp_opt"] = g
        g.replay()

    def optimizer_step(self, norm_from_slots: bool = False) -> None:
        check(self.engine.lib.fnd_clip_adamw_step(self.plan.handle, int(norm_from_slots), self.engine.stream_ptr()),
              "fnd_clip_adamw_step")

    def eval_step(self, from_cache: bool = False, input_set: int = 0) -> None:
        self._launch("eval_step", from_cache, input_set)
        self.plan.bump_all()

    # ------------------------------------------------------------------ results (zero-copy views of the workspace)
    def logits(self) -> torch.Tensor:
        return self.plan.buffer("logits", torch.float32, (self.batch, 2))

    def probs(self) -> torch.Tensor:
        return self.plan.buffer("probs", torch.float32, (self.batch, 2))

    def fused(self) -> torch.Tensor:
        return self.plan.buffer("fused", torch.float32, (self.batch, self.engine.dims.hidden))

    def loss_rows(self) -> torch.Tensor:
        return self.plan.buffer("loss_row", torch.float32, (self.batch,))

    def forensic(self) -> Dict[str, torch.Tensor]:
        rs = self.plan.buffer("rowstat", torch.float32, (self.batch, 16))
        return {"semantic_conflict": rs[:, 0], "emotion_intensity": rs[:, 1], "temporal_delay": rs[:, 2]}

    def loss_scalar_view(self) -> torch.Tensor:
        """Device view of DevState.loss (mean loss of the last training step)."""
        return self.plan.buffer("state", torch.float32, (22,))[13:14]

    def enable_loss_mirror(self, ring: int = 4096) -> torch.Tensor:
        """Per-step losses without a D2H copy: returns a pinned host tensor of `ring` floats; every later training step stores
        its mean loss into slot (optimizer steps taken so far) % ring straight from the device (fnd_set_loss_mirror). Read it
        after the step has completed (stream / event synchronisation) — the host-side counterpart of the reference's
        ``loss.item()`` (forensic_trainer.py:301) without stalling the stream between two steps."""
        self._loss_mirror = torch.zeros(int(ring), dtype=torch.float32).pin_memory()
        check(self.engine.lib.fnd_set_loss_mirror(self.plan.handle, self._loss_mirror.data_ptr(), int(ring), self.engine.stream_ptr()),
              "fnd_set_loss_mirror")
        return self._loss_mirror

    def mark_params_updated(self) -> None:
        """The library's AdamW updates the arena (and the shadows) behind torch's back; keep the modules' shadow
        bookkeeping consistent so a later module-level forward does not refresh needlessly."""
        self.engine._shadow_version = self.engine.param_version()
