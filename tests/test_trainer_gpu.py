"""GPU test of the drop-in ForensicTrainer (fused CUDA-graph epoch loops, device-resident cache, best.pt schema)."""
import os

import numpy as np
import pytest
import torch

from ultrafnd_git_b200.trainer import ForensicTrainer, TrainConfig, synthetic_cache

pytestmark = pytest.mark.gpu


def test_fit_and_test_on_synthetic_cache(tmp_path):
    cache = synthetic_cache(n=600, seed=3)
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir=str(tmp_path), batch_size=64, epochs=4, lr=5e-4)
    tr = ForensicTrainer(cfg, cache=cache)
    assert tuple(tr.cache["gnn_Z"].shape) == (600, 128)
    l0, m0 = tr._epoch_loop("val")
    best = tr.fit()
    l1, m1 = tr._epoch_loop("val")
    print(f"val loss {l0:.4f} -> {l1:.4f}; val auc {m0['auc']:.3f} -> {m1['auc']:.3f}; best {best:.3f}")
    assert l1 < l0 and best > 0.9
    ck = torch.load(os.path.join(str(tmp_path), "best.pt"), map_location="cpu")
    assert set(ck.keys()) == {"fusion", "clf", "gnn", "cfg"}
    assert "fuse_mlp.0.weight" in ck["fusion"] and "node.trees.5.leaf_logits" in ck["clf"] and ck["cfg"]["batch_size"] == 64
    res = tr.test()
    assert set(res) == {"test_loss", "test_acc", "test_auc", "test_precision", "test_recall", "test_f1", "test_cmcs", "test_dfdr"}
    assert res["test_auc"] > 0.85
    # the ragged last batch (420 = 6*64 + 36 train rows) went through its own plan with a consistent optimizer step count
    assert sorted(tr._steps) == [26, 36, 64] or len(tr._steps) >= 2
    steps = {b: s.plan.state()["step"] for b, s in tr._steps.items()}
    print("optimizer step counters per plan:", steps)


def test_forward_batch_api_matches_fused_eval():
    cache = synthetic_cache(n=128, seed=4)
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir="/tmp/fnd_out_api", batch_size=16, epochs=0)
    tr = ForensicTrainer(cfg, cache=cache)
    tr.fusion.eval(); tr.clf.eval()
    batch = next(iter(tr.test_loader))
    with torch.no_grad():
        out = tr._forward_batch(batch, "test")
    assert out["logits"].shape == (len(batch["label"]), 2) and set(out["forensic"]) == {"emotion_intensity", "semantic_conflict", "temporal_delay"}
    st = tr._step_for(len(batch["label"]))
    st.static_gather.copy_(torch.as_tensor(tr.te_idx)[batch["index"]].cuda())
    st.eval_step(from_cache=True)
    assert torch.allclose(st.logits(), out["logits"], atol=1e-5, rtol=1e-4)


def test_resume_continues_bit_for_bit(tmp_path):
    """save_resume / load_resume: two epochs + save + one epoch == load into a fresh trainer + one epoch,
    bitwise (same dropout salts, optimizer step count, Adam moments, shuffle order)."""
    cache = synthetic_cache(n=300, seed=5)
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir=str(tmp_path), batch_size=32, epochs=0, lr=3e-4)
    a = ForensicTrainer(cfg, cache=cache)
    for ep in (1, 2):
        a.epoch = ep
        a._epoch_loop("train")
    path = os.path.join(str(tmp_path), "resume.pt")
    a.save_resume(path)
    a.epoch = 3
    la, _ = a._epoch_loop("train")
    pa = a.engine.params.clone()
    b = ForensicTrainer(cfg, cache=cache)
    b.load_resume(path)
    assert b.epoch == 2
    b.epoch = 3
    lb, _ = b._epoch_loop("train")
    assert la == lb
    assert torch.equal(pa, b.engine.params)
    assert a._last_step.plan.state()["step"] == b._last_step.plan.state()["step"] > 0
