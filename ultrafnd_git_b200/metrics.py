"""Per-epoch metrics with the same keys and definitions as the reference's
``src/training/metrics/forensic_metrics.py`` (:62-171): accuracy, auc, precision, recall, f1 at a 0.5 threshold on the
positive-class probability, CMCS = 1 - mean(clip(0.5*(conflict + delay), 0, 1)), DFDR = TPR of the positive class,
and the mean emotion intensity. Pure numpy (the AUC is the tie-aware rank statistic, equal to sklearn's roc_auc_score).
Host-side, once per epoch — not part of the device hot path."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def _auc(y_true: np.ndarray, y_prob: np.ndarray) -> float:
    y_true = np.asarray(y_true).astype(int)
    y_prob = np.asarray(y_prob).astype(float)
    if y_true.size == 0 or np.unique(y_true).size < 2 or not np.all(np.isfinite(y_prob)):
        return 0.5                                   # forensic_metrics.py:19-33: chance level instead of raising
    order = np.argsort(y_prob, kind="mergesort")
    ranks = np.empty(y_prob.size, dtype=float)
    sorted_p = y_prob[order]
    i = 0
    while i < sorted_p.size:                          # average ranks over ties
        j = i
        while j + 1 < sorted_p.size and sorted_p[j + 1] == sorted_p[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    n_pos = float((y_true == 1).sum())
    n_neg = float(y_true.size - n_pos)
    return float((ranks[y_true == 1].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def compute_classification_metrics(y_true, y_prob, threshold: float = 0.5) -> Dict[str, float]:
    y_true = np.asarray(y_true).astype(int)
    y_prob = np.asarray(y_prob).astype(float)
    y_pred = (y_prob >= threshold).astype(int)
    if y_true.size == 0:
        return {"accuracy": 0.0, "auc": 0.5, "precision": 0.0, "recall": 0.0, "f1": 0.0}
    tp = float(((y_pred == 1) & (y_true == 1)).sum())
    fp = float(((y_pred == 1) & (y_true == 0)).sum())
    fn = float(((y_pred == 0) & (y_true == 1)).sum())
    prec = tp / (tp + fp) if tp + fp > 0 else 0.0
    rec = tp / (tp + fn) if tp + fn > 0 else 0.0
    f1 = 2 * prec * rec / (prec + rec) if prec + rec > 0 else 0.0
    return {"accuracy": float((y_pred == y_true).mean()), "auc": _auc(y_true, y_prob), "precision": prec,
            "recall": rec, "f1": f1}


def compute_cmcs(semantic_conflict, temporal_delay) -> float:
    mix = np.clip(0.5 * (np.asarray(semantic_conflict, float) + np.asarray(temporal_delay, float)), 0.0, 1.0)
    return float(1.0 - mix.mean()) if mix.size else 0.0


def compute_dfdr(y_true, y_prob, threshold: float = 0.5) -> float:
    y_true = np.asarray(y_true).astype(int)
    pos = y_true == 1
    if pos.sum() < 1:
        return 0.0
    return float(((np.asarray(y_prob, float) >= threshold)[pos]).sum() / pos.sum())


def aggregate_epoch_metrics(y_true, y_score, forensic: Optional[Dict[str, np.ndarray]] = None,
                            threshold: float = 0.5) -> Dict[str, float]:
    out = compute_classification_metrics(y_true, y_score, threshold)
    if forensic:
        sc, td = forensic.get("semantic_conflict"), forensic.get("temporal_delay")
        if sc is not None and td is not None:
            out["cmcs"] = compute_cmcs(sc, td)
        ei = forensic.get("emotion_intensity")
        if ei is not None:
            ei = np.asarray(ei, float)
            out["emotion_intensity_mean"] = float(ei.mean()) if ei.size else 0.0
        out["dfdr"] = compute_dfdr(y_true, y_score, threshold)
    return out


def pretty_print(split: str, m: Dict[str, float]) -> None:
    ordered = ["accuracy", "auc", "precision", "recall", "f1", "cmcs", "dfdr"]
    extras = [k for k in m if k not in ordered and not k.startswith("cm_")]
    line = " | ".join(f"{k}:{m[k]:.4f}" for k in ordered if k in m)
    if extras:
        line += " | " + " ".join(f"{k}:{m[k]:.4f}" for k in extras)
    print(f"[{split}] {line}")
