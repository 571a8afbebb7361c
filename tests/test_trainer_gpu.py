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
    l0, m0 = tr._epoch_loop(tr.val_loader, "val")
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


def test_fit_after_load_resume_continues_the_run(tmp_path):
    """fit() after load_resume() must continue with epoch+1 (same shuffle seeds, the StepLR value of an
    uninterrupted run, early-stopping counters kept) — three epochs in one go == two epochs, save, fresh process state,
    load, fit to three."""
    cache = synthetic_cache(n=300, seed=6)
    mk = lambda d, e: TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir=os.path.join(str(tmp_path), d),
                                  batch_size=32, epochs=e, lr=3e-4)
    a = ForensicTrainer(mk("a", 3), cache=cache)
    a.fit()
    b = ForensicTrainer(mk("b", 2), cache=cache)
    b.fit()
    path = os.path.join(str(tmp_path), "resume.pt")
    b.save_resume(path)
    c = ForensicTrainer(mk("c", 3), cache=cache)
    c.load_resume(path)
    assert c.epoch == 2
    c.fit()
    assert c.epoch == 3 and abs(c.lr - 3e-4 * 0.7) < 1e-12 and abs(a.lr - c.lr) < 1e-12
    assert torch.equal(a.engine.params, c.engine.params)
    assert a.best_val_auc == c.best_val_auc and a.no_improve == c.no_improve


def test_epoch_loss_is_mean_of_batch_means():
    """forensic_trainer.py:301,316: the epoch loss is np.mean of the per-batch mean losses (differs from the per-row mean
    when the last batch is short); _epoch_loop keeps the reference's (loader, split) signature."""
    cache = synthetic_cache(n=200, seed=7)        # val split = 30 rows; batch 16 -> batches of 16 and 14
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir="/tmp/fnd_out_bm", batch_size=16, epochs=0)
    tr = ForensicTrainer(cfg, cache=cache)
    loss, m = tr._epoch_loop(tr.val_loader, "val")
    tr.fusion.eval(); tr.clf.eval()
    per_batch = []
    with torch.no_grad():
        for batch in tr.val_loader:
            out = tr._forward_batch(batch, "val")
            per_batch.append(float(torch.nn.functional.cross_entropy(out["logits"], out["y"])))
    assert len(per_batch) == 2
    assert abs(loss - float(np.mean(per_batch))) < 1e-5
    assert tr._epoch_loop("val")[0] == loss


def test_paired_modules_backward_matches_fused_step():
    """trainer._forward_batch(...) followed by loss.backward() (module-level autograd over the PAIRED
    modules) must work and give the fused step's gradients."""
    cache = synthetic_cache(n=128, seed=8)
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir="/tmp/fnd_out_pb", batch_size=16, epochs=0)
    tr = ForensicTrainer(cfg, cache=cache, precision="fp32")
    for m in list(tr.fusion.modules()) + list(tr.clf.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    tr.fusion.train(); tr.clf.train(); tr.fusion._sync_dropout(); tr.clf._sync_dropout()
    batch = next(iter(tr.val_loader))
    out = tr._forward_batch(batch, "val")
    loss = torch.nn.functional.cross_entropy(out["logits"], out["y"])
    loss.backward()
    g_mod = {n: p.grad.clone() for n, p in list(tr.fusion.named_parameters()) + list(tr.clf.named_parameters())
             if p.grad is not None}
    assert "fuse_mlp.0.weight" in g_mod and "pre.0.weight" in g_mod
    st = tr._step_for(len(batch["label"]))
    st.static_gather.copy_(torch.as_tensor(tr.va_idx)[batch["index"]].cuda())
    st.train_fwd_bwd(from_cache=True)
    eng = tr.engine
    worst = 0.0
    for prefix, mod in (("fusion.", tr.fusion), ("clf.", tr.clf)):
        for n, p in mod.named_parameters():
            if p.grad is None:
                continue
            ref = eng.grad_view(prefix + n)
            err = float((p.grad - ref).norm() / (ref.norm() + 1e-12))
            worst = max(worst, err)
    assert abs(float(loss.detach()) - st.plan.state()["loss"]) < 1e-5
    assert worst < 1e-4, worst


def test_epochs_match_reference_trainer():
    """two training epochs + validation + test of the drop-in trainer against the REFERENCE's own
    ForensicTrainer (its real __init__ / _forward_batch / _epoch_loop / metrics on the same synthetic cache; fixture
    tests/golden/trainer_epochs_n300.npz from tests/golden/make_golden.py trainer): per-epoch mean-of-batch-means losses
    and every scalar metric of aggregate_epoch_metrics. fp32 mode, dropout off, same per-epoch sample order, the
    reference's gnn_Z table injected. Losses to 1e-3 relative (north_star fp32 tolerance); count-based metrics may move by
    one borderline sample."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainer_epochs_n300.npz"))
    from oracle import fnd_oracle as O
    cache = synthetic_cache(n=int(g["n"]), seed=int(g["seed"]))
    cache["gnn_Z"] = torch.from_numpy(g["gnn_Z"])
    cfg = TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir="/tmp/fnd_out_epochs", batch_size=int(g["batch_size"]),
                      epochs=2, lr=2e-4, weight_decay=1e-4, seed=42, use_gnn=True, save_best=False)
    tr = ForensicTrainer(cfg, cache=cache, precision="fp32")
    fus_p, clf_p = O.init_params(42)
    O.perturb_node_head(clf_p)
    tr.fusion.load_state_dict(fus_p); tr.clf.load_state_dict(clf_p)
    for m in list(tr.fusion.modules()) + list(tr.clf.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    tr.fusion._sync_dropout(); tr.clf._sync_dropout()
    tr.engine.refresh_shadows(tr.engine.param_version())
    for ep in (1, 2):
        tr.epoch = ep
        tl, tm = tr._epoch_loop(tr.train_loader, "train")
        vl, vm = tr._epoch_loop(tr.val_loader, "val")
        ref_t, ref_v = float(g[f"ep{ep}.train_loss"]), float(g[f"ep{ep}.val_loss"])
        print(f"epoch {ep}: train loss {tl:.6f} (reference {ref_t:.6f}), val loss {vl:.6f} (reference {ref_v:.6f})")
        assert abs(tl - ref_t) <= 1e-3 * ref_t and abs(vl - ref_v) <= 1e-3 * ref_v
        for split, m in (("train", tm), ("val", vm)):
            for k, v in m.items():
                key = f"ep{ep}.{split}.{k}"
                if key in g.files and np.isscalar(v):
                    assert abs(float(v) - float(g[key])) <= 0.02 + 1e-3 * abs(float(g[key])), (key, v, float(g[key]))
    sl, sm = tr._epoch_loop(tr.test_loader, "test")
    assert abs(sl - float(g["test_loss"])) <= 1e-3 * float(g["test_loss"])
    for k, v in sm.items():
        if f"test.{k}" in g.files and np.isscalar(v):
            assert abs(float(v) - float(g[f"test.{k}"])) <= 0.02 + 1e-3 * abs(float(g[f"test.{k}"]))


def test_loss_mirror_delivers_every_steps_loss_to_pinned_host_memory():
    """fnd_set_loss_mirror: the step stores its mean loss (forensic_trainer.py:287 / :301 loss.item()) into a pinned host ring,
    slot = optimizer steps taken so far; it must equal DevState.loss of that step bit for bit, also from a CUDA graph and
    after the ring wraps."""
    import torch
    from oracle import fnd_oracle as O
    from ultrafnd_git_b200.fused import FusedStep
    from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
    torch.manual_seed(9)
    f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
    f.train(); c.train()
    st = FusedStep(f, c, 32, precision="bf16", use_graph=True)
    st.load_batch({k: v.cuda() for k, v in O.make_batch(32, seed=2).items()})
    ring = st.enable_loss_mirror(4)
    seen = []
    for i in range(7):                                    # wraps the 4-slot ring
        st.train_step()
        s = st.plan.state()                               # synchronises
        assert s["step"] == i + 1
        seen.append(s["loss"])
        assert float(ring[i % 4]) == s["loss"], (i, float(ring[i % 4]), s["loss"])
    assert len(set(seen)) > 1
