#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference checkout does not exist on the GPU box):

    HF_HUB_OFFLINE=1 PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights are NOT stored (51 MB): they are regenerated deterministically by ``oracle.fnd_oracle.init_params(seed)``
and loaded into the reference modules with ``load_state_dict``; each fixture stores per-tensor checksums of the
weights so RNG drift is detected. Inputs are stored in full (small batches). Outputs stored: fused, logits,
probs, forensic scalars, loss, per-parameter gradient norms, strided gradient samples, and (train case) the
loss trajectory and strided parameter samples after 3 steps of clip_grad_norm_(5.0) + torch.optim.AdamW.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("FND_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
os.chdir(REF)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import fnd_oracle as O  # noqa: E402
from src.models.fusion.cross_modal_transformer import CrossModalTransformer  # noqa: E402
from src.models.fusion.deep_truth_classifier import DeepTruthClassifier  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STRIDE = 997


def build_reference(seed, perturb):
    fus_p, clf_p = O.init_params(seed)
    if perturb:
        O.perturb_node_head(clf_p)
    fusion = CrossModalTransformer("configs/model_configs/fusion.yaml")
    clf = DeepTruthClassifier("configs/model_configs/classifier.yaml")
    missing = fusion.load_state_dict(fus_p, strict=True)
    clf.load_state_dict(clf_p, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return fusion, clf, fus_p, clf_p


def set_dropout(mod, p):
    for m in mod.modules():
        if isinstance(m, nn.Dropout):
            m.p = p


def checksums(params):
    return {k: np.array([float(v.double().sum()), float((v.double() ** 2).sum())]) for k, v in params.items()}


def ref_forward(fusion, clf, batch):
    feats = {k: batch[k] for k in O.FEAT_KEYS}
    fo = fusion(feats)
    co = clf(fo["fused"], batch["aux"])
    loss = F.cross_entropy(co["logits"], batch["label"])
    return fo, co, loss


def save_case(name, seed, perturb, batch, dist, steps=0):
    torch.manual_seed(0)
    fusion, clf, fus_p, clf_p = build_reference(seed, perturb)
    b = O.make_batch(batch, seed=1234, dist=dist)
    rec = {"meta_seed": np.array(seed), "meta_perturb": np.array(int(perturb)), "meta_batch": np.array(batch),
           "meta_steps": np.array(steps)}
    for k, v in b.items():
        rec["in." + k] = v.numpy()
    for k, v in checksums(fus_p).items():
        rec["wsum.fusion." + k] = v
    for k, v in checksums(clf_p).items():
        rec["wsum.clf." + k] = v

    # ---- eval-mode forward ----
    fusion.eval(); clf.eval()
    with torch.no_grad():
        fo, co, loss = ref_forward(fusion, clf, b)
    rec["eval.fused"] = fo["fused"].numpy()
    rec["eval.fusion_logits"] = fo["logits"].numpy()
    for k, v in fo["forensic"].items():
        rec["eval.forensic." + k] = v.numpy()
    rec["eval.logits"] = co["logits"].numpy()
    rec["eval.probs"] = co["probs"].numpy()
    rec["eval.loss"] = np.array(float(loss))

    # ---- train-mode (dropout forced to 0) gradients ----
    fusion.train(); clf.train()
    set_dropout(fusion, 0.0); set_dropout(clf, 0.0)
    params = list(fusion.parameters()) + list(clf.parameters())
    for p in params:
        p.grad = None
    fo, co, loss = ref_forward(fusion, clf, b)
    loss.backward()
    rec["train.loss"] = np.array(float(loss))
    for prefix, mod in (("fusion", fusion), ("clf", clf)):
        for k, p in mod.named_parameters():
            if p.grad is None:
                rec[f"gnone.{prefix}.{k}"] = np.array(1)
                continue
            g = p.grad.detach()
            rec[f"gnorm.{prefix}.{k}"] = np.array(float(g.double().norm()))
            flat = g.flatten()
            rec[f"gsamp.{prefix}.{k}"] = (flat if flat.numel() <= 4096 else flat[::STRIDE]).numpy().copy()

    # ---- k optimizer steps exactly as forensic_trainer.py:286-298 (dropout off) ----
    if steps:
        opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-4)
        losses, norms = [], []
        for s in range(steps):
            bs = O.make_batch(batch, seed=2000 + s, dist=dist)
            for k, v in bs.items():
                rec[f"step{s}.in.{k}"] = v.numpy()
            fo, co, loss = ref_forward(fusion, clf, bs)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            n = nn.utils.clip_grad_norm_(params, max_norm=5.0)
            opt.step()
            losses.append(float(loss)); norms.append(float(n))
        rec["steps.loss"] = np.array(losses)
        rec["steps.grad_norm"] = np.array(norms)
        for prefix, mod in (("fusion", fusion), ("clf", clf)):
            for k, p in mod.named_parameters():
                flat = p.detach().flatten()
                rec[f"psamp.{prefix}.{k}"] = (flat if flat.numel() <= 4096 else flat[::STRIDE]).numpy().copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB, eval loss {float(rec['eval.loss']):.6f}")


def save_large(name, seed, batch, data_seed):
    """BASELINE.json configs[1]/[2] at full size (train batch 128, eval batch 1024): inputs are NOT stored (7 MB) — they
    are regenerated from ``O.make_batch(batch, seed=data_seed)`` — only the reference's outputs are."""
    torch.manual_seed(0)
    fusion, clf, fus_p, clf_p = build_reference(seed, True)
    b = O.make_batch(batch, seed=data_seed)
    rec = {"meta_seed": np.array(seed), "meta_batch": np.array(batch), "meta_data_seed": np.array(data_seed),
           "in_checksum": np.array([float(sum(v.double().sum() for v in b.values()))])}
    fusion.eval(); clf.eval()
    with torch.no_grad():
        fo, co, loss = ref_forward(fusion, clf, b)
    rec["eval.logits"] = co["logits"].numpy()
    rec["eval.probs"] = co["probs"].numpy()
    rec["eval.fused_rowsum"] = fo["fused"].double().sum(-1).numpy()
    rec["eval.loss"] = np.array(float(loss))
    for k, v in fo["forensic"].items():
        rec["eval.forensic." + k] = v.numpy()
    fusion.train(); clf.train()
    set_dropout(fusion, 0.0); set_dropout(clf, 0.0)
    params = list(fusion.parameters()) + list(clf.parameters())
    fo, co, loss = ref_forward(fusion, clf, b)
    loss.backward()
    rec["train.loss"] = np.array(float(loss))
    rec["train.grad_norm"] = np.array(float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params if p.grad is not None))))
    for prefix, mod in (("fusion", fusion), ("clf", clf)):
        for k, p in mod.named_parameters():
            if p.grad is not None:
                rec[f"gnorm.{prefix}.{k}"] = np.array(float(p.grad.double().norm()))
    os.makedirs(os.path.join(OUT, "large"), exist_ok=True)
    path = os.path.join(OUT, "large", name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB, eval loss {float(rec['eval.loss']):.6f}")


from tests.golden.gcn_common import gcn_inputs  # noqa: E402  (shared with the tests)


def save_gcn(name, n, seed):
    """the reference's OWN SimpleGCN / build_adj_from_ocr / pre-train loop (forensic_trainer.py:25-53,
    113-132,184-224) on a synthetic post graph, dropout 0 so the result is deterministic (the reference draws its
    dropout(0.2) masks from torch's CPU RNG, which no other implementation can replay). Inputs and initial weights are
    regenerated from seeds by the test; only a strided sample of gnn_Z, its row norms and the adjacency degrees are stored."""
    from src.training.forensic_trainer import SimpleGCN, build_adj_from_ocr
    X, ocr = gcn_inputs(n, seed)
    adj = build_adj_from_ocr(ocr, thresh=0.12)
    torch.manual_seed(seed)
    gnn = SimpleGCN(in_dim=416, hid=256, out_dim=128, dropout=0.0)
    head = nn.Linear(128, 1)
    init = {k: v.detach().clone() for k, v in list(gnn.state_dict().items())}
    head_init = {k: v.detach().clone() for k, v in head.state_dict().items()}
    Xt, At = torch.from_numpy(X), torch.from_numpy(adj)
    opt = torch.optim.Adam(gnn.parameters(), lr=1e-3, weight_decay=1e-4)          # _pretrain_gnn, forensic_trainer.py:213-224
    target = At.sum(dim=-1, keepdim=True) / max(1.0, At.shape[0])
    losses = []
    for _ in range(2):
        gnn.train()
        Z = gnn(Xt, At)
        loss = F.mse_loss(torch.sigmoid(head(Z)), target)
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
    with torch.no_grad():
        Z = gnn(Xt, At)
    out = {"n": np.array(n), "seed": np.array(seed), "degree": adj.sum(-1).astype(np.float32), "edges": np.array(adj.sum()),
           "pretrain_losses": np.array(losses), "Z_rows": Z[::16].numpy(), "Z_row_norm": Z.norm(dim=1).numpy(),
           "Z_fro": np.array(float(Z.double().norm()))}
    for k, v in init.items():
        out["init." + k] = v.numpy()
    for k, v in head_init.items():
        out["head." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: n={n} edges={int(adj.sum())} max degree={int(adj.sum(-1).max())} pretrain losses {losses} |Z|={float(Z.norm()):.4f}")


def save_trainer_epochs(name, n=300, seed=9, batch_size=32, epochs=2):
    """the reference's REAL ForensicTrainer (its __init__, _build_gnn, _forward_batch, _epoch_loop and
    metrics code, forensic_trainer.py:139-330) on a synthetic feature cache. Only the out-of-scope data pipeline is
    replaced (FakeSVRawDataset / build_gnn_cache_from_raw_dataset return the synthetic cache). Deterministic set-up:
    weights from O.init_params(42) (+ perturbed NODE head), every dropout p = 0, and the train loader iterates in the
    seeded per-epoch permutation the drop-in trainer uses (the reference's shuffle=True draws from the global CPU RNG).
    Stored: the reference's gnn_Z table (it carries the reference's dropout draw, so the test injects it), per-epoch
    train / val (loss, metrics) and the optimizer trajectory's final loss."""
    import src.training.forensic_trainer as RT
    from ultrafnd_git_b200.trainer import synthetic_cache
    cache = synthetic_cache(n=n, seed=seed)
    RT.FakeSVRawDataset = lambda root: None
    RT.build_gnn_cache_from_raw_dataset = lambda raw, **kw: cache
    outdir = "/tmp/fnd_golden_trainer"
    cfg = RT.TrainConfig(data_root="unused", ocr_phrase_pkl=None, out_dir=outdir, batch_size=batch_size, epochs=epochs, lr=2e-4,
                         weight_decay=1e-4, seed=42, use_mps=False, use_gnn=True, save_best=False)
    tr = RT.ForensicTrainer(cfg)
    fus_p, clf_p = O.init_params(42)
    O.perturb_node_head(clf_p)
    tr.fusion.load_state_dict(fus_p, strict=True)
    tr.clf.load_state_dict(clf_p, strict=True)
    for m in (tr.fusion, tr.clf, tr.gnn):
        set_dropout(m, 0.0)
    out = {"n": np.array(n), "seed": np.array(seed), "batch_size": np.array(batch_size), "gnn_Z": tr.cache["gnn_Z"].numpy()}
    tr_ds = tr.train_loader.dataset
    for ep in range(1, epochs + 1):
        g = torch.Generator().manual_seed(cfg.seed * 7919 + ep)
        order = torch.randperm(len(tr_ds), generator=g).tolist()
        loader = torch.utils.data.DataLoader(tr_ds, batch_size=batch_size, sampler=order, drop_last=False)
        tl, tm = tr._epoch_loop(loader, "train")
        vl, vm = tr._epoch_loop(tr.val_loader, "val")
        out[f"ep{ep}.train_loss"] = np.array(tl); out[f"ep{ep}.val_loss"] = np.array(vl)
        for k, v in tm.items():
            if np.isscalar(v):
                out[f"ep{ep}.train.{k}"] = np.array(float(v))
        for k, v in vm.items():
            if np.isscalar(v):
                out[f"ep{ep}.val.{k}"] = np.array(float(v))
        print(f"{name}: epoch {ep} train loss {tl:.6f} val loss {vl:.6f} val auc {vm.get('auc')}")
    tl, tm = tr._epoch_loop(tr.test_loader, "test")
    out["test_loss"] = np.array(tl)
    for k, v in tm.items():
        if np.isscalar(v):
            out[f"test.{k}"] = np.array(float(v))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "trainer":
        save_trainer_epochs("trainer_epochs_n300")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gcn":
        save_gcn("gcn_n512", 512, 21)
        save_gcn("gcn_n5504", 5504, 22)
        sys.exit(0)
    save_case("eval_smoke_b4", seed=42, perturb=False, batch=4, dist="smoke")
    save_case("trained_cache_b16", seed=42, perturb=True, batch=16, dist="cache")
    save_case("train3_smoke_b8", seed=43, perturb=True, batch=8, dist="smoke", steps=3)
    save_large("train_b128", seed=42, batch=128, data_seed=31)
    save_large("eval_b1024", seed=42, batch=1024, data_seed=9)
