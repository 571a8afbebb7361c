"""B200 drop-in for the reference's live trainer (src/training/forensic_trainer.py): same ``TrainConfig`` fields,
``ForensicTrainer(cfg)`` with ``fit() -> best_val_auc`` and ``test() -> dict``, the ``best.pt`` schema
``{"fusion", "clf", "gnn", "cfg"}`` (:352-361), StepLR(3, 0.7) per epoch (:177,341), patience-3 early stopping on
validation AUC (:362-366), and ``_forward_batch`` / ``_build_dataloaders`` / ``*_loader`` for the callers that poke at them
(scripts/sanity_check.py:25-26).

What changes underneath:
  * the step is ``fnd_train_step`` / ``fnd_eval_step`` (one CUDA-graph replay per batch) instead of ~150 ATen ops
    + autograd + a Python optimizer loop;
  * the whole feature cache and the GCN embeddings live on the device as one matrix; a batch is a gather by row index
    inside the first kernel (no DataLoader / collate / .to(device) per step);
  * the six blocking D2H copies per step (:301-313) are replaced by device-side accumulation and ONE copy per epoch;
  * with ``torch.distributed`` initialised, each global batch is sharded across ranks; with the NCCL backend the
    optimizer step is ``fnd_train_step_dp`` — gradient reduce-scatter out of NVLink peer memory, sharded AdamW, bf16
    shadow all-gather (csrc/fnd_dp.cuh) — and ``FND_DP=nccl`` selects a plain gradient all-reduce between
    ``fnd_train_fwd_bwd`` and ``fnd_clip_adamw_step`` instead.
The data pipeline and graph construction stay the reference's (out of scope, SURVEY.md §2 #9,#14): pass a prebuilt
``cache`` dict, or have the reference importable as ``src.data_pipeline.fakesv_dataset``.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import check
from .fused import DeviceCache, FusedStep, FEATURE_KEYS
from .gcn import SimpleGCN, build_gnn_embeddings
from .metrics import aggregate_epoch_metrics, pretty_print
from .modules import CrossModalTransformer, DeepTruthClassifier, pair_modules


@dataclass
class TrainConfig:
    """Field-for-field the reference's TrainConfig (forensic_trainer.py:90-107)."""
    data_root: str
    ocr_phrase_pkl: Optional[str]
    out_dir: str = "outputs"
    batch_size: int = 16
    epochs: int = 8
    lr: float = 2e-4
    weight_decay: float = 1e-4
    gnn_dim: int = 128
    gnn_overlap_thresh: float = 0.12
    seed: int = 42
    use_mps: bool = True          # accepted and ignored: this trainer runs on CUDA
    use_gnn: bool = True
    save_best: bool = True
    grad_clip: float = 5.0
    early_stop_patience: int = 3


def build_adj_from_ocr(ocr_sets, thresh: float = 0.12) -> np.ndarray:
    """Jaccard-threshold adjacency (forensic_trainer.py:113-132) computed from a token-incidence matrix instead of an
    O(N^2) Python loop; identical result (1 on the diagonal; edge iff |a∩b| / (|a∪b| + 1e-9) >= thresh)."""
    n = len(ocr_sets)
    vocab: Dict[str, int] = {}
    rows, cols = [], []
    for i, s in enumerate(ocr_sets):
        for tok in s:
            rows.append(i)
            cols.append(vocab.setdefault(tok, len(vocab)))
    inc = torch.zeros(n, max(len(vocab), 1), dtype=torch.float32)
    if rows:
        inc[torch.tensor(rows), torch.tensor(cols)] = 1.0
    inter = inc @ inc.t()
    size = inc.sum(-1)
    union = size[:, None] + size[None, :] - inter
    jac = inter / (union + 1e-9)
    both_empty = (size[:, None] == 0) & (size[None, :] == 0)
    a = ((jac >= thresh) & ~both_empty).float()
    a.fill_diagonal_(1.0)
    return a.numpy()


def shard_indices(global_idx: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's share of one global batch: every world-th sample starting at `rank` (disjoint, exhaustive; shard
    sizes differ by at most one). With per-rank loss scale 1/len(global batch) a SUM all-reduce of the gradients yields
    exactly the single-process mean-loss gradient."""
    return global_idx[rank::world] if world > 1 else global_idx


def gather_ragged(tensors: List[torch.Tensor], count: int, world: int, group=None) -> Tuple[List[torch.Tensor], int]:
    """All-gather the first ``count`` rows of every tensor from every rank when the counts differ between ranks: the
    counts are exchanged first, shards are padded to the longest one (the buffers must have at least that many rows)
    and the padding is cut away again. Returns (rank-major concatenations, total rows); identical on every rank."""
    dev = tensors[0].device
    cnt = torch.tensor([count], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    torch.distributed.all_gather(cnts, cnt, group=group)
    cnts = [int(c.item()) for c in cnts]
    m = max(max(cnts), 1)
    out = []
    for t in tensors:
        if t.shape[0] < m:
            t = torch.cat([t, t.new_zeros((m - t.shape[0],) + tuple(t.shape[1:]))])
        mine = t[:m].contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(parts, mine, group=group)
        out.append(torch.cat([p[:c] for p, c in zip(parts, cnts)]))
    return out, sum(cnts)


def mean_of_batch_means(loss_rows: torch.Tensor, batch_id: torch.Tensor, nbatches: int) -> float:
    """np.mean over batches of the per-batch mean loss (forensic_trainer.py:301,316) from per-row losses and the global
    batch number of every row; differs from the per-row mean whenever the last batch is short."""
    sums = torch.zeros(nbatches, device=loss_rows.device, dtype=torch.float64).index_add_(0, batch_id, loss_rows.double())
    cnt = torch.zeros(nbatches, device=loss_rows.device, dtype=torch.float64).index_add_(
        0, batch_id, torch.ones_like(loss_rows, dtype=torch.float64))
    seen = cnt > 0
    return float((sums[seen] / cnt[seen]).mean().cpu())


class CachedTensorDataset(torch.utils.data.Dataset):
    """API-compatibility view of one split (forensic_trainer.py:60-83); the fused loops do not use it."""

    def __init__(self, cache: Dict, indices: np.ndarray):
        self.T = torch.from_numpy(np.asarray(cache["text"])[indices])
        self.A = torch.from_numpy(np.asarray(cache["audio"])[indices])
        self.V = torch.from_numpy(np.asarray(cache["visual"])[indices])
        self.U = torch.from_numpy(np.asarray(cache["temporal"])[indices])
        self.AUX = torch.from_numpy(np.asarray(cache["aux"])[indices])
        self.y = torch.from_numpy(np.asarray(cache["labels"])[indices]).long()

    def __len__(self):
        return self.T.shape[0]

    def __getitem__(self, i):
        return {"text_features": self.T[i], "audio_features": self.A[i], "visual_features": self.V[i],
                "temporal_features": self.U[i], "aux": self.AUX[i], "label": self.y[i], "index": i}


class ForensicTrainer:
    def __init__(self, cfg: TrainConfig, cache: Optional[Dict] = None, precision: Optional[str] = None,
                 use_graph: bool = True):
        self.cfg = cfg
        os.makedirs(cfg.out_dir, exist_ok=True)
        if not torch.cuda.is_available():
            raise RuntimeError("ultrafnd_git_b200.ForensicTrainer needs a CUDA device (no CPU fallback)")
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world = torch.distributed.get_world_size() if self.dist else 1
        self.rank = torch.distributed.get_rank() if self.dist else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.float32
        torch.manual_seed(cfg.seed)
        np.random.seed(cfg.seed)

        if cache is None:
            try:
                from src.data_pipeline.fakesv_dataset import FakeSVRawDataset, build_gnn_cache_from_raw_dataset
            except Exception as e:  # pragma: no cover - needs the reference checkout
                raise RuntimeError("no feature cache given and the reference data pipeline "
                                   "(src.data_pipeline.fakesv_dataset) is not importable") from e
            raw = FakeSVRawDataset(cfg.data_root)
            cache = build_gnn_cache_from_raw_dataset(raw, ocr_phrase_pkl=cfg.ocr_phrase_pkl, text_dim=768, audio_dim=128,
                                                     visual_dim=512, temporal_dim=256, seed=cfg.seed)
        self.cache = dict(cache)         # shallow copy: _build_gnn adds "gnn_Z" without mutating the caller's dict
        cache = self.cache
        self.tr_idx, self.va_idx, self.te_idx = (np.asarray(x) for x in cache["split"])
        if "gnn_Z" not in cache or cache["gnn_Z"] is None:
            self._build_gnn()
        else:
            self.gnn = None
        self.train_loader, self.val_loader, self.test_loader = self._build_dataloaders()

        self.fusion = CrossModalTransformer(config_path="configs/model_configs/fusion.yaml", precision=precision)
        self.clf = DeepTruthClassifier(config_path="configs/model_configs/classifier.yaml", precision=precision)
        if cfg.use_gnn is False and self.fusion.use_gnn:
            # the reference crashes in fuse_mlp.0 in this configuration (SURVEY.md §7 hard parts); say so up front
            raise RuntimeError("use_gnn=False with fusion.yaml use_gnn: true: the reference's fuse_mlp.0 is sized for 16 "
                               "slots and fails with a shape error; edit fusion.yaml as well")
        self.engine = pair_modules(self.fusion, self.clf, precision)
        # Data parallel: identical replicas (rank 0's arena is broadcast), batches sharded by rank. With the NCCL backend
        # the optimizer step runs sharded over NVLink peer memory (csrc/fnd_dp.cuh; FND_DP=nccl selects the plain
        # gradient all-reduce + replicated AdamW instead).
        self.dp_peer = False
        if self.dist:
            torch.distributed.broadcast(self.engine.params, src=0)
            if torch.distributed.get_backend() == "nccl" and os.environ.get("FND_DP", "peer") == "peer":
                self.engine.enable_optimizer()
                self.engine.enable_symmetric(torch.distributed.group.WORLD)
                self.dp_peer = True
        self.engine.set_hyper(lr=cfg.lr, weight_decay=cfg.weight_decay, max_norm=float(cfg.grad_clip or 0.0))
        self.engine.set_seed(cfg.seed + 1000003 * self.rank)
        self.use_graph = use_graph
        self.precision = precision
        self._steps: Dict[int, FusedStep] = {}
        self._last_step: Optional[FusedStep] = None
        self._idx_stream: Optional[torch.cuda.Stream] = None     # stages the next batch's row indices (_epoch_loop)
        self._resume_state: Optional[torch.Tensor] = None
        self._resumed = False

        feats = {"text_features": torch.as_tensor(np.asarray(cache["text"])),
                 "audio_features": torch.as_tensor(np.asarray(cache["audio"])),
                 "visual_features": torch.as_tensor(np.asarray(cache["visual"])),
                 "temporal_features": torch.as_tensor(np.asarray(cache["temporal"])),
                 "gnn_feat": torch.as_tensor(np.asarray(torch.as_tensor(cache["gnn_Z"]).detach().cpu()))}
        self.dcache = DeviceCache(feats, torch.as_tensor(np.asarray(cache["aux"])),
                                  torch.as_tensor(np.asarray(cache["labels"])), self.device)
        self.lr = cfg.lr
        self.epoch = 0
        self.best_val_auc = -1.0
        self.no_improve = 0
        self.ckpt_path = os.path.join(cfg.out_dir, "best.pt")

    # ------------------------------------------------------------------ start-up stages (not the hot path)
    def _build_gnn(self) -> None:
        """forensic_trainer.py:184-224: compact node features, OCR-Jaccard graph, 2-epoch degree-regression pre-train,
        then a constant embedding table — on the library's tensor-core GEMM (gcn.py: one adjacency GEMM per Â·x, Â never
        materialised; no cuBLAS). FND_GNN_Z_DROPOUT=1 reproduces the reference's quirk of leaving the GCN in train mode
        (dropout 0.2 active) while the cached table is computed; the default computes it deterministically."""
        c, cfg = self.cache, self.cfg
        T, A, V, U = (np.asarray(c[k]) for k in ("text", "audio", "visual", "temporal"))
        X = np.concatenate([T[:, :192], A[:, :32], V[:, :128], U[:, :64]], axis=1).astype(np.float32)
        X /= (np.linalg.norm(X, axis=1, keepdims=True) + 1e-9)
        adj = build_adj_from_ocr(c["ocr_sets"], thresh=cfg.gnn_overlap_thresh)
        self.X = torch.from_numpy(X).to(self.device)
        self.Adj = torch.from_numpy(adj).to(self.device)
        self.gnn = SimpleGCN(self.X.shape[1], hid=2 * cfg.gnn_dim, out_dim=cfg.gnn_dim, dropout=0.2).to(self.device)
        head = nn.Linear(cfg.gnn_dim, 1).to(self.device)          # constructed even when unused: same RNG consumption
        self.cache["gnn_Z"] = build_gnn_embeddings(self.X, self.Adj, self.gnn, head, pretrain=bool(cfg.use_gnn), epochs=2,
                                                   z_dropout=os.environ.get("FND_GNN_Z_DROPOUT", "0") == "1")
        self.gnn.eval()

    def _build_dataloaders(self):
        mk = lambda idx, shuf: torch.utils.data.DataLoader(CachedTensorDataset(self.cache, idx),
                                                            batch_size=self.cfg.batch_size, shuffle=shuf, drop_last=False)
        return mk(self.tr_idx, True), mk(self.va_idx, False), mk(self.te_idx, False)

    # ------------------------------------------------------------------ module-level forward (API compatibility)
    def _forward_batch(self, batch, split: str) -> Dict[str, torch.Tensor]:
        """forensic_trainer.py:238-271 through the module API (autograd-capable); the epoch loops use the fused path."""
        idx = {"train": self.tr_idx, "val": self.va_idx}.get(split, self.te_idx)
        gidx = torch.as_tensor(idx, device=self.device)[batch["index"].to(self.device)]
        feats = {k: batch[k].to(self.device, dtype=self.dtype) for k in FEATURE_KEYS[:4]}
        feats["gnn_feat"] = torch.as_tensor(self.cache["gnn_Z"]).to(self.device)[gidx] if self.cfg.use_gnn else None
        fo = self.fusion(feats)
        co = self.clf(fo["fused"], batch["aux"].to(self.device, dtype=self.dtype))
        return {"logits": co["logits"], "probs": co["probs"], "y": batch["label"].to(self.device),
                "forensic": fo.get("forensic", {})}

    # ------------------------------------------------------------------ fused epoch loop
    def _step_for(self, batch: int, loss_scale: Optional[float] = None) -> FusedStep:
        """The fused step for a per-rank batch size. ``loss_scale`` = d(mean loss)/d(row loss): 1/len(global batch) under
        data parallelism (a SUM of the rank gradients is then the single-process gradient), 0 for a padding step (a rank
        whose shard of a short last batch is empty still takes part in the collective optimizer step, contributing zero)."""
        st = self._steps.get(batch)
        if st is None:
            st = FusedStep(self.fusion, self.clf, batch, precision=self.precision, use_graph=self.use_graph)
            st.attach_cache(self.dcache)
            st._loss_scale = 1.0 / batch
            self._steps[batch] = st
        if loss_scale is not None and st._loss_scale != loss_scale:
            from ._lib import check
            check(st.engine.lib.fnd_set_loss_scale(st.plan.handle, float(loss_scale), st.engine.stream_ptr()),
                  "fnd_set_loss_scale")
            st._loss_scale = loss_scale
        if self._last_step is None and self._resume_state is not None:
            # resuming: the saved DevState (optimizer step, bias corrections, dropout salts, hyper-parameters) seeds
            # the first plan that runs
            dst = st.plan.buffer("state", torch.float32, (22,))
            keep_scale = dst[17:18].clone()
            dst.copy_(self._resume_state.to(self.device))
            dst[17:18].copy_(keep_scale)
            dst[16:17].zero_(); dst[18:20].zero_()        # error flag, election / intra-launch counters
            self._resume_state = None
        if self._last_step is not None and self._last_step is not st:
            # optimizer step count / bias corrections / dropout salts live in each plan's DevState: carry them over
            src = self._last_step.plan.buffer("state", torch.float32, (22,))
            dst = st.plan.buffer("state", torch.float32, (22,))
            keep_scale = dst[17:18].clone()
            dst.copy_(src)
            dst[17:18].copy_(keep_scale)
        self._last_step = st
        return st

    def _epoch_loop(self, loader, split: Optional[str] = None) -> Tuple[float, Dict[str, float]]:
        """forensic_trainer.py:273-330 with the reference's signature ``_epoch_loop(loader, split)``. The loader argument
        is accepted for call compatibility; batches are gathered on the device from the resident feature cache in the
        loader's order for val/test (sequential) and in a seeded per-epoch permutation for train (the reference's
        ``shuffle=True`` draws from torch's global CPU RNG, which cannot be reproduced bit-for-bit anyway).
        ``_epoch_loop("train")`` (split only) is accepted too.
        Returns (mean of the per-batch mean losses — forensic_trainer.py:301,316 —, metrics over the whole split);
        under data parallelism every rank returns the same global values."""
        if split is None:
            split = loader
        is_train = split == "train"
        idx_np = {"train": self.tr_idx, "val": self.va_idx}.get(split, self.te_idx)
        n = len(idx_np)
        self.fusion.train(is_train); self.clf.train(is_train)
        order = torch.as_tensor(idx_np, dtype=torch.int64)
        if is_train:
            g = torch.Generator().manual_seed(self.cfg.seed * 7919 + self.epoch)
            order = order[torch.randperm(n, generator=g)]
        order = order.to(self.device)
        bs = self.cfg.batch_size
        # device-side epoch buffers: one D2H copy at the end instead of six per step (forensic_trainer.py:301-313)
        cap = (n + self.world - 1) // self.world + (n + bs - 1) // max(bs, 1) + 1
        loss_rows = torch.zeros(cap, device=self.device)
        p1 = torch.zeros(cap, device=self.device)
        ys = torch.zeros(cap, dtype=torch.int64, device=self.device)
        bid = torch.zeros(cap, dtype=torch.int64, device=self.device)       # global batch number of every row
        forensic = torch.zeros(cap, 3, device=self.device)
        done = 0
        nbatches = 0
        # Row indices of batch i+1 are staged into the second of the step's two index buffers by a side stream while the graph
        # of batch i runs: a small copy enqueued in front of every graph launch delays the step's first kernel (measured on
        # the bench: 5 us per step for the 1 KB index copy, 15 us for a 4-byte read-back).
        starts = list(range(0, n, bs))
        main = torch.cuda.current_stream(self.device)
        if self._idx_stream is None:
            self._idx_stream = torch.cuda.Stream(self.device)
        side = self._idx_stream
        side.wait_stream(main)                               # `order` was produced on the main stream
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in free:
            e.record(main)

        def stage(bi: int):
            gidx = order[starts[bi]:starts[bi] + bs]
            local = shard_indices(gidx, self.rank, self.world)
            k = int(local.numel())
            if k == 0 and not (self.dist and is_train):
                return gidx, local, k
            # (the loss scale is set when the step runs, not here: it may differ between two batches of one size)
            st = self._step_for(max(k, 1))
            with torch.cuda.stream(side):
                side.wait_event(free[bi % 2])                # the graph that last read this buffer has finished
                st.gather_set(bi % 2).copy_(local if k else gidx[:1], non_blocking=True)
                ready[bi % 2].record(side)
            return gidx, local, k

        staged = stage(0) if starts else None
        for bno in range(len(starts)):
            gidx, local, k = staged
            nbatches += 1
            s = bno % 2
            if k == 0:
                # short last batch (the reference keeps it: drop_last=False) with fewer samples than ranks: this rank has
                # nothing to evaluate, but a data-parallel optimizer step is collective, so it runs one padding row with
                # loss weight 0 (zero gradient contribution) instead of leaving the other ranks waiting
                if not (self.dist and is_train):
                    staged = stage(bno + 1) if bno + 1 < len(starts) else None
                    continue
                st = self._step_for(1, 0.0)
            else:
                st = self._step_for(k, 1.0 / int(gidx.numel()) if self.dist else None)
            main.wait_event(ready[s])
            if is_train:
                if self.dp_peer:
                    st.train_step_dp(from_cache=True, input_set=s)
                elif self.dist:
                    st.train_fwd_bwd(from_cache=True, input_set=s)
                    torch.distributed.all_reduce(st.engine.grads)
                    st.optimizer_step(norm_from_slots=False)
                else:
                    st.train_step(from_cache=True, input_set=s)
            else:
                st.eval_step(from_cache=True, input_set=s)
            free[s].record(main)
            staged = stage(bno + 1) if bno + 1 < len(starts) else None
            if k == 0:
                continue
            # one launch appends this step's rows (row losses, p1, the labels the step gathered, forensic scalars, batch number)
            check(st.engine.lib.fnd_collect_rows(st.plan.handle, k, done, bno, loss_rows.data_ptr(), p1.data_ptr(), ys.data_ptr(),
                                                 bid.data_ptr(), forensic.data_ptr(), st.engine.stream_ptr()), "fnd_collect_rows")
            done += k
        if is_train and self._last_step is not None:
            self._last_step.mark_params_updated()
            self._last_step.plan.check_error()
        if self.dist:
            # ALWAYS gather (shard sizes may differ by a few rows), so that every rank computes the same loss / AUC and
            # takes the same checkpoint and early-stopping decisions
            (loss_rows, p1, ys, bid, forensic), done = gather_ragged([loss_rows, p1, ys, bid, forensic], done, self.world)
        else:
            loss_rows, p1, ys, bid, forensic = loss_rows[:done], p1[:done], ys[:done], bid[:done], forensic[:done]
        loss_mean = mean_of_batch_means(loss_rows, bid, nbatches) if done else 0.0
        f = forensic.cpu().numpy()
        metrics = aggregate_epoch_metrics(ys.cpu().numpy(), p1.cpu().numpy(),
                                          {"semantic_conflict": f[:, 0], "emotion_intensity": f[:, 1],
                                           "temporal_delay": f[:, 2]} if done else None, threshold=0.5)
        return loss_mean, metrics

    def fit(self) -> float:
        """forensic_trainer.py:332-369. After ``load_resume`` it continues with the epoch after the saved one (same
        shuffle seeds, learning-rate schedule and early-stopping counters as an uninterrupted run)."""
        if not self._resumed:
            self.no_improve = 0
        self._resumed = False
        for epoch in range(self.epoch + 1, self.cfg.epochs + 1):
            self.epoch = epoch
            tr_loss, tr_m = self._epoch_loop(self.train_loader, "train")
            va_loss, va_m = self._epoch_loop(self.val_loader, "val")
            # StepLR(step_size=3, gamma=0.7), stepped once per epoch (forensic_trainer.py:177,341); closed form so that a
            # resumed run lands on the same value
            lr = self.cfg.lr * (0.7 ** (epoch // 3))
            if lr != self.lr:
                self.lr = lr
                self.engine.set_lr(self.lr)
            if self.rank == 0:
                print(f"[Epoch {epoch:02d}] train_loss={tr_loss:.4f} | ", end=""); pretty_print("train", tr_m)
                print(f"           val_loss={va_loss:.4f} | ", end=""); pretty_print("val", va_m)
            val_auc = float(va_m.get("auc", 0.5))
            if self.dist:
                # identical on every rank by construction (same gathered rows); rank 0's value is authoritative so that a
                # last-bit difference can never split the ranks between "save" (a collective) and "stop"
                t = torch.tensor([val_auc], dtype=torch.float64, device=self.device)
                torch.distributed.broadcast(t, src=0)
                val_auc = float(t.item())
            if val_auc > self.best_val_auc + 1e-4 and self.cfg.save_best:
                self.best_val_auc = val_auc
                self.no_improve = 0
                if self.dp_peer:
                    self.engine.gather_master()      # fp32 master weights are sharded between optimizer steps
                if self.rank == 0:
                    tmp = self.ckpt_path + ".tmp"
                    torch.save({"fusion": {k: v.detach().cpu() for k, v in self.fusion.state_dict().items()},
                                "clf": {k: v.detach().cpu() for k, v in self.clf.state_dict().items()},
                                "gnn": self.gnn.state_dict() if (self.cfg.use_gnn and self.gnn is not None) else None,
                                "cfg": dict(self.cfg.__dict__)}, tmp)
                    os.replace(tmp, self.ckpt_path)  # readers never see a partial file
                    print(f"  ↳ saved best checkpoint to {self.ckpt_path} (val_auc={self.best_val_auc:.3f})")
            else:
                self.no_improve += 1
                if self.no_improve >= self.cfg.early_stop_patience:
                    if self.rank == 0:
                        print(f"↳ Early stopping (no val AUC improvement for {self.cfg.early_stop_patience} epochs)")
                    break
        return self.best_val_auc

    # ------------------------------------------------------------------ full resume
    def save_resume(self, path: str) -> None:
        """Everything needed to continue training bit-for-bit: fp32 master weights, Adam moments, the device-side step
        state (optimizer step count, dropout salts), learning rate, epoch and early-stopping bookkeeping. The reference
        only saves the best weights (forensic_trainer.py:352-361) and cannot resume. Every rank must call this (the
        sharded data-parallel state is gathered first); rank 0 writes."""
        eng = self.engine
        if self.dp_peer:
            eng.gather_master()
            for r in range(self.world):                      # Adam moments are sharded like the master weights
                for lo, hi in eng.shard_ranges(r):
                    if hi > lo:
                        torch.distributed.broadcast(eng.adam_m[lo:hi], src=r)
                        torch.distributed.broadcast(eng.adam_v[lo:hi], src=r)
        else:
            eng.dp_flush()
        dev_state = (self._last_step.plan.buffer("state", torch.float32, (22,)).cpu().clone()
                     if self._last_step is not None else self._resume_state)
        if self.rank == 0:
            torch.save({"params": eng.params.detach().cpu(),
                        "adam_m": eng.adam_m.cpu() if eng.adam_m is not None else None,
                        "adam_v": eng.adam_v.cpu() if eng.adam_v is not None else None,
                        "dev_state": dev_state, "lr": self.lr, "epoch": self.epoch, "best_val_auc": self.best_val_auc,
                        "no_improve": self.no_improve, "cfg": dict(self.cfg.__dict__)}, path)

    def load_resume(self, path: str) -> None:
        ck = torch.load(path, map_location="cpu")
        eng = self.engine
        eng.enable_optimizer()
        with torch.no_grad():
            eng.params.copy_(ck["params"].to(self.device))
            if ck.get("adam_m") is not None:
                eng.adam_m.copy_(ck["adam_m"].to(self.device))
                eng.adam_v.copy_(ck["adam_v"].to(self.device))
        eng.refresh_shadows(eng.param_version())
        self.lr = float(ck["lr"]); eng.set_lr(self.lr)
        self.epoch = int(ck["epoch"]); self.best_val_auc = float(ck["best_val_auc"]); self.no_improve = int(ck["no_improve"])
        self._resume_state = ck["dev_state"]
        self._resumed = True
        self._steps.clear()
        self._last_step = None

    def test(self) -> Dict[str, float]:
        if self.dist:
            torch.distributed.barrier()          # rank 0's best.pt (written + renamed in fit) is complete before anyone reads
        if os.path.exists(self.ckpt_path):
            ck = torch.load(self.ckpt_path, map_location="cpu")
            self.fusion.load_state_dict(ck["fusion"])
            self.clf.load_state_dict(ck["clf"])
            if self.cfg.use_gnn and ck.get("gnn") is not None and self.gnn is not None:
                self.gnn.load_state_dict(ck["gnn"])
            self.engine.refresh_shadows(self.engine.param_version())
        ts_loss, m = self._epoch_loop(self.test_loader, "test")
        if self.rank == 0:
            print(f"[Test] loss={ts_loss:.4f} | ", end=""); pretty_print("test", m)
        return {"test_loss": ts_loss, "test_acc": m.get("accuracy", 0.0), "test_auc": m.get("auc", 0.5),
                "test_precision": m.get("precision", 0.0), "test_recall": m.get("recall", 0.0), "test_f1": m.get("f1", 0.0),
                "test_cmcs": m.get("cmcs", 0.0), "test_dfdr": m.get("dfdr", 0.0)}


def synthetic_cache(n: int = 256, seed: int = 0) -> Dict:
    """A FakeSV-shaped feature cache with the keys build_gnn_cache_from_raw_dataset returns (fakesv_dataset.py:242-252),
    filled with separable synthetic data (tests, smoke runs, benchmarks: there is no dataset offline)."""
    g = np.random.RandomState(seed)
    labels = g.randint(0, 2, size=n).astype(np.int64)
    shift = (labels[:, None] * 2 - 1).astype(np.float32)

    def feat(d, scale):
        x = g.randn(n, d).astype(np.float32)
        x[:, : d // 8] += scale * shift
        return x / (np.linalg.norm(x, axis=1, keepdims=True) + 1e-9)
    idx = g.permutation(n)
    n_tr, n_va = int(0.7 * n), int(0.15 * n)
    vocab = [f"tok{i}" for i in range(64)]
    ocr = [set(g.choice(vocab, size=g.randint(0, 6), replace=False).tolist()) for _ in range(n)]
    return {"ids": np.array([f"v{i}" for i in range(n)]), "text": feat(768, 0.6), "audio": feat(128, 0.4),
            "visual": feat(512, 0.5), "temporal": (0.05 * g.randn(n, 256)).astype(np.float32),
            "aux": g.rand(n, 2).astype(np.float32), "labels": labels, "ocr_sets": ocr,
            "split": (idx[:n_tr], idx[n_tr:n_tr + n_va], idx[n_tr + n_va:])}
