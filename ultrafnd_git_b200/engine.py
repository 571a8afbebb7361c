"""Host-side engine: the flat parameter arena and the per-batch plans, bound to libfnd_b200.so through ctypes.

PyTorch is plumbing here: it owns device memory (the arena, gradient / optimizer buffers, bf16 shadows and each
plan's workspace are plain ``torch`` tensors whose ``data_ptr()`` is handed to the C ABI) and supplies the CUDA
stream. All arithmetic of the hot path runs in the library's kernels. There is no CPU or eager fallback: any
compute entry point raises when CUDA or the library is unavailable.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import FndDims, FndInputs, check

MODE_BF16 = 0
MODE_FP32X3 = 1

LAYER_STREAMS = {"fuse0": 1, "fuse1": 2, "pre0": 3, "pre1": 4, "tree": 5}


def default_precision() -> int:
    """bf16 unless FND_PRECISION=fp32 (the bf16x3 tensor-core mode with fp32-equivalent results)."""
    v = os.environ.get("FND_PRECISION", "bf16").lower()
    if v in ("bf16", "0"):
        return MODE_BF16
    if v in ("fp32", "fp32x3", "1"):
        return MODE_FP32X3
    raise ValueError(f"FND_PRECISION={v!r}: expected 'bf16' or 'fp32'")


@dataclass
class Dims:
    """Model dimensions (fusion.yaml + classifier.yaml of the reference; input widths are fixed by
    cross_modal_transformer.py:96-99)."""
    hidden: int = 512
    d_text: int = 768
    d_audio: int = 128
    d_visual: int = 512
    d_temporal: int = 256
    d_gnn: int = 128
    use_gnn: bool = True
    aux_dim: int = 2
    trees: int = 6
    depth: int = 4
    fusion_dropout: float = 0.1
    clf_dropout: float = 0.1
    tree_dropout: float = 0.3
    node_tau: float = 10.0

    def to_c(self) -> FndDims:
        return FndDims(self.hidden, self.d_text, self.d_audio, self.d_visual, self.d_temporal, self.d_gnn,
                       int(self.use_gnn), self.aux_dim, self.trees, self.depth, self.fusion_dropout,
                       self.clf_dropout, self.tree_dropout, self.node_tau)

    def validate(self) -> None:
        if self.hidden not in (512, 1024):
            raise NotImplementedError(f"hidden_dim={self.hidden}: the sm_100a kernels support 512 and 1024")
        if self.aux_dim not in (0, 2):
            raise NotImplementedError(f"aux_dim={self.aux_dim}: supported values are 0 and 2")
        if self.depth > 4 or self.trees * self.depth > 32 or self.trees < 1 or self.depth < 1:
            raise NotImplementedError("NODE head: depth <= 4 and trees*depth <= 32 are supported")
        if self.d_gnn % 64:
            raise NotImplementedError("gnn_dim must be a multiple of 64")


@dataclass
class ParamInfo:
    name: str
    offset: int
    shape: Tuple[int, ...]
    hot: bool

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def param_table(dims: Dims) -> List[ParamInfo]:
    """The arena layout as the library defines it (fnd_param_info)."""
    lib = _lib.load()
    cd = dims.to_c()
    n = lib.fnd_param_count(ctypes.byref(cd))
    if n <= 0:
        raise _lib.FndError(f"fnd_param_count rejected dims {dims}")
    out = []
    name = ctypes.create_string_buffer(128)
    off = ctypes.c_longlong()
    ndim, rows, cols, hot = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for i in range(n):
        check(lib.fnd_param_info(ctypes.byref(cd), i, name, 128, ctypes.byref(off), ctypes.byref(ndim),
                                 ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hot)), "fnd_param_info")
        shape: Tuple[int, ...] = () if ndim.value == 0 else ((rows.value,) if ndim.value == 1 else (rows.value, cols.value))
        out.append(ParamInfo(name.value.decode(), off.value, shape, bool(hot.value)))
    return out


class Plan:
    """One (batch, mode) plan: workspace tensor + the library's plan handle."""

    def __init__(self, engine: "Engine", batch: int):
        self.engine = engine
        self.batch = batch
        self.lib = engine.lib
        handle = ctypes.c_void_p()
        cd = engine.dims.to_c()
        check(self.lib.fnd_plan_create(ctypes.byref(cd), batch, engine.mode, ctypes.byref(handle)), "fnd_plan_create")
        self.handle = handle
        nbytes = self.lib.fnd_plan_workspace_bytes(handle)
        self.workspace = torch.zeros(nbytes + 256, dtype=torch.uint8, device=engine.device)
        self._ws_base = (self.workspace.data_ptr() + 255) // 256 * 256
        self._ws_shift = self._ws_base - self.workspace.data_ptr()
        # Activation generations: the fusion stage and the classifier stage save their activations in disjoint workspace
        # buffers, so each has its own counter (a paired classifier forward must not invalidate the fusion forward that
        # fed it). The fused entry points overwrite both.
        self.fusion_id = 0
        self.clf_id = 0
        self.bind()

    @property
    def forward_id(self) -> int:
        return self.fusion_id + self.clf_id

    def bump_all(self) -> None:
        self.fusion_id += 1
        self.clf_id += 1

    def bind(self) -> None:
        e = self.engine
        check(self.lib.fnd_plan_set_grad_mirror(self.handle, e.grads_bf.data_ptr() if e.grads_bf is not None else None),
              "fnd_plan_set_grad_mirror")
        check(self.lib.fnd_plan_bind(self.handle, self._ws_base, e.params.data_ptr(), e.grads.data_ptr(),
                                     e.adam_m.data_ptr() if e.adam_m is not None else None,
                                     e.adam_v.data_ptr() if e.adam_v is not None else None,
                                     e.shadow_hi.data_ptr(), e.shadow_lo.data_ptr() if e.shadow_lo is not None else None,
                                     e.stream_ptr()), "fnd_plan_bind")
        h = e.hyper
        check(self.lib.fnd_set_hyper(self.handle, h["lr"], h["beta1"], h["beta2"], h["eps"], h["weight_decay"],
                                     h["max_norm"], e.stream_ptr()), "fnd_set_hyper")
        check(self.lib.fnd_set_seed(self.handle, e.seed, e.stream_ptr()), "fnd_set_seed")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.fnd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def buffer(self, name: str, dtype: torch.dtype, shape: Tuple[int, ...]) -> torch.Tensor:
        """Zero-copy view of a named workspace buffer."""
        off = self.lib.fnd_plan_buffer_offset(self.handle, name.encode())
        if off < 0:
            raise KeyError(name)
        nbytes = self.lib.fnd_plan_buffer_bytes(self.handle, name.encode())
        raw = self.workspace[self._ws_shift + off: self._ws_shift + off + nbytes]
        t = raw.view(dtype)
        n = 1
        for s in shape:
            n *= s
        return t[:n].view(*shape)

    def state(self) -> Dict[str, float]:
        """Reads DevState (synchronises)."""
        raw = self.buffer("state", torch.uint8, (88,)).cpu().numpy().tobytes()
        import struct
        rng = struct.unpack_from("<4I", raw, 0)
        lr, b1, b2, eps, wd, mx, bc1, bc2 = struct.unpack_from("<8f", raw, 16)
        step, = struct.unpack_from("<i", raw, 48)
        loss, gnorm, coef = struct.unpack_from("<3f", raw, 52)
        err, = struct.unpack_from("<i", raw, 64)
        lscale, = struct.unpack_from("<f", raw, 68)
        return {"rng": rng, "lr": lr, "beta1": b1, "beta2": b2, "eps": eps, "weight_decay": wd, "max_norm": mx,
                "bc1": bc1, "bc2": bc2, "step": step, "loss": loss, "grad_norm": gnorm, "clip_coef": coef,
                "err": err, "loss_scale": lscale}

    def check_error(self) -> None:
        check(self.lib.fnd_check_error(self.handle, self.engine.stream_ptr()), "device error flag")

    def launch_count(self, entry: str) -> int:
        return self.lib.fnd_launch_count(self.handle, entry.encode())

    def dropout_masks(self) -> Dict[str, torch.Tensor]:
        """Keep-multipliers the NEXT training forward will draw (test support: replayed in the CPU oracle)."""
        d, B = self.engine.dims, self.batch
        shapes = {"fuse0": (B, 2 * d.hidden), "fuse1": (B, d.hidden), "pre0": (B, d.hidden), "pre1": (B, d.hidden),
                  "tree": (B, d.trees, 2)}
        out = {}
        for k, shp in shapes.items():
            n = 1
            for s in shp:
                n *= s
            t = torch.empty(n, dtype=torch.float32, device=self.engine.device)
            check(self.lib.fnd_export_dropout_mask(self.handle, LAYER_STREAMS[k], t.data_ptr(), n,
                                                   self.engine.stream_ptr()), "fnd_export_dropout_mask")
            out[k] = t.view(*shp)
        return out


class Engine:
    """Flat fp32 parameter arena (+ grads, Adam moments, bf16 operand shadows) and the plans that run on it."""

    def __init__(self, dims: Dims, device: Optional[torch.device] = None, mode: Optional[int] = None):
        dims.validate()
        self.lib = _lib.load()
        self.dims = dims
        self.mode = default_precision() if mode is None else mode
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self.table = param_table(dims)
        self.index = {p.name: p for p in self.table}
        cd = dims.to_c()
        self.n_total = self.lib.fnd_arena_total_elems(ctypes.byref(cd))
        self.n_hot = self.lib.fnd_arena_hot_elems(ctypes.byref(cd))
        self.n_shadow = self.lib.fnd_arena_shadow_elems(ctypes.byref(cd))
        self.n_shadow_buf = self.lib.fnd_arena_shadow_buffer_elems(ctypes.byref(cd))
        self.params = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
        self.grads: Optional[torch.Tensor] = None
        self.shadow_hi: Optional[torch.Tensor] = None
        self.shadow_lo: Optional[torch.Tensor] = None
        self.adam_m: Optional[torch.Tensor] = None
        self.adam_v: Optional[torch.Tensor] = None
        self.plans: Dict[int, Plan] = {}
        self.hyper = {"lr": 2e-4, "beta1": 0.9, "beta2": 0.999, "eps": 1e-8, "weight_decay": 1e-4, "max_norm": 5.0}
        self.seed = 0x5EED5EED
        self._shadow_version: Optional[int] = None
        self.attached: list = []          # weakrefs to the nn.Modules whose parameters live in this arena
        self.symm: Optional[dict] = None  # peer-mapped buffers of the data-parallel optimizer step (enable_symmetric)
        self.grads_bf: Optional[torch.Tensor] = None   # bf16 mirror of the GEMM-weight gradients (pull-mode exchange)
        self._dp_pending = False          # a deferred fuse_mlp update of the last train_step_dp has not been applied yet
        self._dp_plan: Optional["Plan"] = None
        self._alloc_device_buffers()

    def param_version(self) -> int:
        """Sum of the autograd version counters of every attached parameter: changes whenever torch code (an
        optimizer step, load_state_dict, ...) writes a parameter in place, which is when the bf16 shadows go stale."""
        total = 0
        for ref in self.attached:
            m = ref()
            if m is not None and getattr(m, "_engine", None) is self:
                total += m._param_version()
        return total

    # ---------------------------------------------------------------- memory
    def _alloc_device_buffers(self) -> None:
        if self.device.type != "cuda":
            return
        self.grads = torch.zeros(self.n_hot, dtype=torch.float32, device=self.device)
        self.shadow_hi = torch.zeros(self.n_shadow_buf, dtype=torch.bfloat16, device=self.device)
        self.shadow_lo = torch.zeros(self.n_shadow_buf, dtype=torch.bfloat16, device=self.device) if self.mode == MODE_FP32X3 else None

    def enable_optimizer(self) -> None:
        """Allocate the Adam moments (needed by fnd_clip_adamw_step / fnd_train_step) and re-bind the plans."""
        self.require_cuda()
        if self.adam_m is None:
            self.adam_m = torch.zeros(self.n_hot, dtype=torch.float32, device=self.device)
            self.adam_v = torch.zeros(self.n_hot, dtype=torch.float32, device=self.device)
            for p in self.plans.values():
                p.bind()

    # ---------------------------------------------------------------- data parallel over peer memory
    def enable_symmetric(self, group=None) -> None:
        """Move params / grads / bf16 shadows into ONE peer-mapped (symmetric) allocation shared with the other ranks
        of ``group`` and bind the sharded optimizer step (csrc/fnd_dp.cuh). torch.distributed's symmetric-memory
        allocator is plumbing here: it hands out the CUDA VMM mapping of every peer's buffer; all data movement is done
        by this library's kernels. Call once, after the modules are attached and before any plan is used for training.
        Raises if peer mapping is unavailable — callers that want an all-reduce path must choose it explicitly."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.require_cuda()
        self.enable_optimizer()
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            raise NotImplementedError("the peer-memory optimizer step supports up to 8 ranks (one NVSwitch domain)")

        def up(x):
            return (x + 255) // 256 * 256
        off_p = 0
        off_g = up(off_p + 4 * self.n_total)
        off_h = up(off_g + 4 * self.n_hot)
        off_l = up(off_h + 2 * self.n_shadow_buf)
        off_pad = up(off_l + (2 * self.n_shadow_buf if self.mode == MODE_FP32X3 else 0))
        off_stage = off_pad + 256
        # gradients travel as bf16 in bf16 mode (summed in fp32 by the owner) unless FND_DP_GRAD=fp32; always fp32 in fp32 mode
        stage_bf16 = int(self.mode == MODE_BF16 and os.environ.get("FND_DP_GRAD", "bf16") != "fp32")
        stage_bytes = self.lib.fnd_dp_stage_bytes(self.any_plan().handle, world, stage_bf16)
        if stage_bytes < 0:
            raise _lib.FndError(f"fnd_dp_stage_bytes: {stage_bytes}")
        # pull mode (NVSwitch multicast): the reduce-scatter is one multimem.ld_reduce kernel; in bf16 mode the GEMM-weight
        # gradients are reduced from a bf16 mirror that the wgrad kernels write (FND_DP_PULL=1; default: store-based exchange)
        # Off by default: on the 8-GPU box the in-switch reduction measured 312.6 us/step against 306.4 us for the
        # store-based exchange (the pull itself is shorter, 61 vs 77 us, but the mirror stores slow the wgrad kernel).
        want_pull = os.environ.get("FND_DP_PULL", "0") == "1"
        off_gbf = up(off_stage + stage_bytes)
        total = off_gbf + (2 * self.n_hot if (want_pull and stage_bf16) else 0)
        buf = symm.empty(total, dtype=torch.uint8, device=self.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        new_p = buf[off_p: off_p + 4 * self.n_total].view(torch.float32)
        new_g = buf[off_g: off_g + 4 * self.n_hot].view(torch.float32)
        new_h = buf[off_h: off_h + 2 * self.n_shadow_buf].view(torch.bfloat16)
        new_l = buf[off_l: off_l + 2 * self.n_shadow_buf].view(torch.bfloat16) if self.mode == MODE_FP32X3 else None
        new_p.copy_(self.params)
        new_h.copy_(self.shadow_hi)
        if new_l is not None:
            new_l.copy_(self.shadow_lo)
        self.params, self.grads, self.shadow_hi, self.shadow_lo = new_p, new_g, new_h, new_l
        multicast = int(getattr(hdl, "multicast_ptr", 0) or 0) if os.environ.get("FND_DP_MULTICAST", "1") != "0" else 0
        pull = bool(multicast) and want_pull
        self.grads_bf = buf[off_gbf: off_gbf + 2 * self.n_hot].view(torch.bfloat16) if (pull and stage_bf16) else None
        with torch.no_grad():
            for ref in self.attached:
                m = ref()
                if m is not None and getattr(m, "_engine", None) is self:
                    for name, p in m.named_parameters():
                        p.data = self.view(m._prefix + name)
        self.plans.clear()
        self._shadow_version = None
        per = stage_bytes // (world * (2 if stage_bf16 else 4))      # elements of the largest slice
        self.symm = {"buf": buf, "handle": hdl, "rank": rank, "world": world, "group": group, "stage_bf16": stage_bf16,
                     "offsets": (off_p, off_g, off_h, off_l, off_pad, off_stage),
                     "peer_bases": [int(x) for x in hdl.buffer_ptrs],
                     # NVSwitch multicast mapping (0 when unsupported, or when disabled with FND_DP_MULTICAST=0)
                     "multicast": multicast,
                     # -2: store-based exchange; -1: pull from the fp32 arenas; >= 0: pull, offset of the bf16 mirror
                     "off_grads_bf16": (off_gbf if stage_bf16 else -1) if pull else -2,
                     "gred": torch.zeros(per, dtype=torch.float32, device=self.device),
                     "slots": torch.zeros(1024, dtype=torch.float32, device=self.device)}
        torch.cuda.synchronize(self.device)
        dist.barrier(group)          # every rank's pad is zeroed and mapped before anyone signals

    def dp_bind(self, plan: "Plan") -> None:
        s = self.symm
        bases = (ctypes.c_ulonglong * s["world"])(*s["peer_bases"])
        off_p, off_g, off_h, off_l, off_pad, off_stage = s["offsets"]
        check(self.lib.fnd_dp_bind(plan.handle, s["rank"], s["world"], bases, off_p, off_g, off_h, off_l, off_pad, off_stage,
                                   s["stage_bf16"], s["multicast"], s["off_grads_bf16"],
                                   s["gred"].data_ptr(), s["gred"].numel(), s["slots"].data_ptr(), s["slots"].numel()),
              "fnd_dp_bind")

    def shard_ranges(self, rank: int) -> List[Tuple[int, int]]:
        lo, hi = (ctypes.c_longlong * 3)(), (ctypes.c_longlong * 3)()
        n = self.lib.fnd_dp_shard_ranges(self.any_plan().handle, rank, self.symm["world"], lo, hi)
        if n < 0:
            raise _lib.FndError(f"fnd_dp_shard_ranges: {n}")
        return [(lo[i], hi[i]) for i in range(n)]

    def dp_flush(self) -> None:
        """Apply a pending deferred data-parallel update now (every rank must call this at the same point)."""
        if self._dp_pending and self._dp_plan is not None:
            check(self.lib.fnd_dp_flush(self._dp_plan.handle, self.stream_ptr()), "fnd_dp_flush")
        self._dp_pending = False

    def gather_master(self) -> None:
        """After sharded optimizer steps only the owner of a slice holds current fp32 master weights: broadcast every
        slice from its owner so that ``state_dict()`` / checkpoints see the full model on every rank."""
        import torch.distributed as dist
        s = getattr(self, "symm", None)
        if s is None:
            return
        self.dp_flush()
        for r in range(s["world"]):
            for lo, hi in self.shard_ranges(r):
                if hi > lo:
                    dist.broadcast(self.params[lo:hi], src=dist.get_global_rank(s["group"], r), group=s["group"])

    def gather_reduced_grads(self) -> torch.Tensor:
        """The REDUCED gradient of the last data-parallel step as one [n_hot] fp32 tensor on every rank (test / bench
        support: after fnd_train_step_dp each rank only holds the reduced values of its own slice, segments back to
        back in ``symm["gred"]``)."""
        import torch.distributed as dist
        s = self.symm
        if s is None:
            raise RuntimeError("enable_symmetric() first")
        out = torch.zeros(self.n_hot, dtype=torch.float32, device=self.device)
        for r in range(s["world"]):
            off = 0
            for lo, hi in self.shard_ranges(r):
                n = hi - lo
                if n > 0:
                    if r == s["rank"]:
                        out[lo:hi].copy_(s["gred"][off:off + n])
                    dist.broadcast(out[lo:hi], src=dist.get_global_rank(s["group"], r), group=s["group"])
                off += n
        return out

    def view(self, name: str) -> torch.Tensor:
        p = self.index[name]
        return self.params[p.offset: p.offset + p.numel].view(p.shape)

    def grad_view(self, name: str, grads: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        p = self.index[name]
        if not p.hot:
            return None
        g = self.grads if grads is None else grads
        return g[p.offset: p.offset + p.numel].view(p.shape)

    def to(self, device: torch.device) -> None:
        device = torch.device(device)
        if device == self.device:
            return
        self.params = self.params.to(device)
        self.device = device
        self.plans.clear()
        self.adam_m = self.adam_v = None
        self.grads = self.shadow_hi = self.shadow_lo = None
        self.grads_bf = None
        self.symm = None
        self._shadow_version = None
        self._alloc_device_buffers()

    def set_mode(self, mode: int) -> None:
        if mode != self.mode:
            self.mode = mode
            self.plans.clear()
            self._shadow_version = None
            self._alloc_device_buffers()
            self.adam_m = self.adam_v = None

    def set_dropout(self, fusion_p: Optional[float] = None, clf_p: Optional[float] = None, tree_p: Optional[float] = None) -> None:
        changed = False
        for attr, v in (("fusion_dropout", fusion_p), ("clf_dropout", clf_p), ("tree_dropout", tree_p)):
            if v is not None and float(v) != getattr(self.dims, attr):
                setattr(self.dims, attr, float(v))
                changed = True
        if changed:
            self.plans.clear()

    # ---------------------------------------------------------------- plumbing
    def require_cuda(self) -> None:
        if self.device.type != "cuda":
            raise RuntimeError("ultrafnd_git_b200 runs on CUDA (sm_100a) only: no CPU fallback exists. "
                               "Move the module to a CUDA device.")

    def stream_ptr(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def plan(self, batch: int) -> Plan:
        self.require_cuda()
        p = self.plans.get(batch)
        if p is None:
            with torch.cuda.device(self.device):
                p = Plan(self, batch)
            self.plans[batch] = p
        return p

    def any_plan(self) -> Plan:
        if not self.plans:
            return self.plan(1)
        return next(iter(self.plans.values()))

    def refresh_shadows(self, version: Optional[int] = None) -> None:
        """Rebuild the bf16 operand copies of the GEMM weights from the fp32 master arena."""
        self.require_cuda()
        check(self.lib.fnd_refresh_shadows(self.any_plan().handle, self.stream_ptr()), "fnd_refresh_shadows")
        self._shadow_version = version

    def ensure_shadows(self, version: int) -> None:
        if self._shadow_version != version:
            self.refresh_shadows(version)

    def set_hyper(self, **kw) -> None:
        self.dp_flush()
        self.hyper.update(kw)
        for p in self.plans.values():
            h = self.hyper
            check(self.lib.fnd_set_hyper(p.handle, h["lr"], h["beta1"], h["beta2"], h["eps"], h["weight_decay"],
                                         h["max_norm"], self.stream_ptr()), "fnd_set_hyper")

    def set_lr(self, lr: float) -> None:
        self.dp_flush()
        self.hyper["lr"] = lr
        for p in self.plans.values():
            check(self.lib.fnd_set_lr(p.handle, lr, self.stream_ptr()), "fnd_set_lr")

    def set_seed(self, seed: int) -> None:
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        for p in self.plans.values():
            check(self.lib.fnd_set_seed(p.handle, self.seed, self.stream_ptr()), "fnd_set_seed")


def make_inputs(feats: Dict[str, Optional[torch.Tensor]], aux: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, gather: Optional[torch.Tensor] = None,
                use_gnn: bool = True) -> Tuple[FndInputs, list]:
    """Pack device tensors into an ``fnd_inputs`` struct. Returns the struct and the tensors it points into (keep
    them alive until the call has been enqueued)."""
    keys = ["text_features", "audio_features", "visual_features", "temporal_features"] + (["gnn_feat"] if use_gnn else [])
    inp = FndInputs()
    keep = []
    for i, k in enumerate(keys):
        t = feats[k]
        if t.dtype != torch.float32 or t.stride(-1) != 1:
            t = t.to(torch.float32).contiguous()
        if t.stride(0) % 4 or t.data_ptr() % 16:
            t = t.contiguous().clone()
        keep.append(t)
        inp.x[i] = t.data_ptr()
        inp.pitch[i] = t.stride(0) if t.shape[0] > 1 else t.shape[-1]
    if aux is not None:
        if aux.dtype != torch.float32 or aux.stride(-1) != 1:
            aux = aux.to(torch.float32).contiguous()
        keep.append(aux)
        inp.aux = aux.data_ptr()
        inp.aux_pitch = aux.stride(0) if aux.shape[0] > 1 else aux.shape[-1]
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
        keep.append(labels)
        inp.labels = labels.data_ptr()
    if gather is not None:
        gather = gather.to(torch.int64).contiguous()
        keep.append(gather)
        inp.gather = gather.data_ptr()
    return inp, keep
