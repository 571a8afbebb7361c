"""GPU parity of the B200 fusion hot path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances are BASELINE.json's: logits/loss rel-err <= 1e-3 in fp32 mode and <= 2e-2 in bf16 mode, argmax identical.
"""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import fnd_oracle as O
from ultrafnd_git_b200.fused import FusedStep
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier, pair_modules

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-3, "bf16": 2e-2}
STRIDE = 997


def build_pair(seed=42, perturb=False, precision="fp32", dropout_off=False):
    fus, clf = O.init_params(seed)
    if perturb:
        O.perturb_node_head(clf)
    f = CrossModalTransformer(precision=precision)
    c = DeepTruthClassifier(precision=precision)
    f.load_state_dict(fus)
    c.load_state_dict(clf)
    if dropout_off:
        for m in list(f.modules()) + list(c.modules()):
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    return f, c, fus, clf


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


def load_gold(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    batch = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in.")}
    return z, batch


def module_forward(f, c, batch):
    fo = f({k: batch[k] for k in O.FEAT_KEYS})
    co = c(fo["fused"], batch["aux"])
    return fo, co


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["eval_smoke_b4", "trained_cache_b16", "train3_smoke_b8"])
def test_eval_forward_matches_golden_and_oracle(name, precision):
    z, batch = load_gold(name)
    f, c, fus, clf = build_pair(int(z["meta_seed"]), bool(int(z["meta_perturb"])), precision)
    f.eval(); c.eval()
    with torch.no_grad():
        fo, co = module_forward(f, c, to_cuda(batch))
        ref = O.model_forward(fus, clf, batch)
    tol = TOL[precision]
    got = {"fused": fo["fused"], "fusion_logits": fo["logits"], "logits": co["logits"], "probs": co["probs"]}
    for k, v in got.items():
        e_or = O.rel_err(v.cpu(), ref[k])
        e_gold = O.rel_err(v.cpu(), torch.from_numpy(z["eval." + k]))
        print(f"[{name}/{precision}] {k}: rel-err vs oracle {e_or:.2e}, vs reference golden {e_gold:.2e}")
        assert e_or < tol and e_gold < tol, k
    for k, v in fo["forensic"].items():
        assert O.rel_err(v.cpu(), torch.from_numpy(z["eval.forensic." + k])) < tol, k
    assert torch.equal(co["logits"].argmax(-1).cpu(), torch.from_numpy(z["eval.logits"]).argmax(-1))
    assert float(co["temperature"]) == 1.0
    loss = torch.nn.functional.cross_entropy(co["logits"].cpu(), batch["label"])
    assert abs(float(loss) - float(z["eval.loss"])) / float(z["eval.loss"]) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["trained_cache_b16", "train3_smoke_b8"])
def test_module_autograd_gradients_match_reference(name, precision):
    """model(batch) -> F.cross_entropy -> loss.backward(), exactly the reference trainer's calls, dropout forced off."""
    z, batch = load_gold(name)
    f, c, fus, clf = build_pair(int(z["meta_seed"]), bool(int(z["meta_perturb"])), precision, dropout_off=True)
    f.train(); c.train()
    cb = to_cuda(batch)
    fo, co = module_forward(f, c, cb)
    loss = torch.nn.functional.cross_entropy(co["logits"], cb["label"])
    loss.backward()
    tol = TOL[precision]
    assert abs(float(loss.detach()) - float(z["train.loss"])) / float(z["train.loss"]) < tol
    _, gf, gc = O.loss_and_grads(fus, clf, batch, dropout=0.0)
    worst = 0.0
    for prefix, mod, grads in (("fusion", f, gf), ("clf", c, gc)):
        for k, p in mod.named_parameters():
            if f"gnone.{prefix}.{k}" in z.files:
                assert p.grad is None, f"{prefix}.{k} must stay grad=None like the reference"
                continue
            assert p.grad is not None, f"{prefix}.{k} got no gradient"
            ref = grads[k]
            rn = float(ref.norm())
            if rn == 0.0:
                assert float(p.grad.abs().max()) < 1e-12, k
                continue
            e = O.rel_err(p.grad.cpu(), ref)
            worst = max(worst, e)
            # golden (reference) strided samples as a second witness
            flat = p.grad.flatten().cpu()
            samp = flat if flat.numel() <= 4096 else flat[::STRIDE]
            eg = O.rel_err(samp, torch.from_numpy(z[f"gsamp.{prefix}.{k}"]))
            # gradients pass through ~6 bf16-rounded layers: the full-tensor bound is 5x the logits tolerance; the
            # strided golden sample (a few dozen elements of the big tensors) is noisier, so it gets 10x
            assert e < 5 * tol and eg < 10 * tol, (prefix, k, e, eg)
    print(f"[{name}/{precision}] worst per-parameter gradient rel-err {worst:.2e}")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_train_steps_match_reference(precision):
    """fnd_train_step x3 (fused fwd + CE + bwd + clip + AdamW, CUDA-graph replayed) vs the reference's 3 real steps."""
    z, _ = load_gold("train3_smoke_b8")
    f, c, fus, clf = build_pair(int(z["meta_seed"]), True, precision, dropout_off=True)
    f.train(); c.train()
    f._sync_dropout(); c._sync_dropout()
    B = int(z["meta_batch"])
    step = FusedStep(f, c, B, precision=precision, use_graph=True)
    assert step.engine.dims.fusion_dropout == 0.0 and step.engine.dims.tree_dropout == 0.0
    losses, norms = [], []
    for s in range(int(z["meta_steps"])):
        batch = {k.split(".in.")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"step{s}.in.")}
        step.load_batch(to_cuda(batch))
        step.train_step()
        st = step.plan.state()
        losses.append(st["loss"]); norms.append(st["grad_norm"])
        assert st["err"] == 0 and st["step"] == s + 1
    step.plan.check_error()
    tol = TOL[precision]
    print(f"[train3/{precision}] losses {losses} vs {z['steps.loss']}; norms {norms} vs {z['steps.grad_norm']}")
    np.testing.assert_allclose(losses, z["steps.loss"], rtol=tol)
    np.testing.assert_allclose(norms, z["steps.grad_norm"], rtol=max(tol, 2e-3))
    # parameters after 3 AdamW steps: every weight moved by ~3*lr, compare the values (update error << lr)
    lr = 2e-4
    for prefix, mod in (("fusion", f), ("clf", c)):
        for k, p in mod.named_parameters():
            key = f"psamp.{prefix}.{k}"
            flat = p.detach().flatten().cpu()
            samp = flat if flat.numel() <= 4096 else flat[::STRIDE]
            ref = torch.from_numpy(z[key])
            diff = (samp - ref).abs()
            # Adam's first updates are ~lr*sign(g): elements whose gradient is ~eps (or flips sign in bf16) may differ
            # by a sizeable fraction of the total move, the mean must not
            max_bound = (0.3 if precision == "fp32" else 2.0) * 3 * lr
            mean_bound = (0.01 if precision == "fp32" else 0.3) * 3 * lr
            assert float(diff.max()) <= max_bound and float(diff.mean()) <= mean_bound, (prefix, k, float(diff.max()), float(diff.mean()))


@pytest.mark.parametrize("precision,B", [("fp32", 16), ("bf16", 128)])
def test_train_mode_with_dropout_matches_oracle_given_same_masks(precision, B):
    """Dropout ON: export the Philox keep-masks the next forward will draw, replay them in the CPU oracle.
    ("bf16", 128) is the BENCHMARKED regime (BASELINE.json configs[1]: bf16 operands, dropout on, batch 128)."""
    f, c, fus, clf = build_pair(42, True, precision)
    f.train(); c.train()
    batch = O.make_batch(B, seed=77)
    step = FusedStep(f, c, B, precision=precision, use_graph=False)
    masks = {k: v.cpu() for k, v in step.plan.dropout_masks().items()}
    for k, p in (("fuse0", 0.1), ("fuse1", 0.1), ("pre0", 0.1), ("pre1", 0.1), ("tree", 0.3)):
        keep = float((masks[k] > 0).float().mean())
        vals = torch.unique(masks[k]).tolist()
        assert all(v == 0.0 or abs(v - 1.0 / (1.0 - p)) < 1e-6 for v in vals), vals
        # keep-rate within five binomial standard deviations of 1 - p for THIS mask's element count (tree masks are small:
        # B x trees x 2), plus the 16-bit threshold's quantisation
        n = masks[k].numel()
        bound = 5.0 * math.sqrt(p * (1.0 - p) / n) + 2.0 / 65536.0
        assert abs(keep - (1.0 - p)) < bound, (k, keep, n, bound)
    step.load_batch(to_cuda(batch))
    step.train_fwd_bwd()
    st = step.plan.state()
    out, gf, gc = O.loss_and_grads(fus, clf, batch, dropout=0.1, masks=masks)
    print(f"[dropout/{precision}/B={B}] loss {st['loss']} vs oracle {float(out['loss'])}")
    assert abs(st["loss"] - float(out["loss"])) / float(out["loss"]) < TOL[precision]
    eng = step.engine
    worst = ("", 0.0)
    total = math.sqrt(sum(float(g.norm()) ** 2 for grads in (gf, gc) for g in grads.values()))
    floor = 1e-4 * total
    for prefix, grads in (("fusion", gf), ("clf", gc)):
        for k, g in grads.items():
            got = eng.grad_view(f"{prefix}.{k}").cpu()
            gn = float(g.norm())
            if gn == 0:
                continue
            if gn < floor:
                # Vanishing gradients (the evidence-gate MLPs at batch 128: |g| ~ 1e-7 .. 1e-5 against a total norm of 0.29,
                # a batch sum that cancels almost completely): a RELATIVE error of such a tensor measures rounding noise —
                # fp32 mode reproduces them to 2e-3, bf16 operands to 8e-2 .. 0.18. They are held to an ABSOLUTE bound at
                # the scale below which a gradient cannot move the step: 5 x tol x 1e-4 of the total gradient norm.
                assert float((got - g).norm()) < 5 * TOL[precision] * floor, (prefix, k, float((got - g).norm()), gn)
                continue
            e = O.rel_err(got, g)
            if e > worst[1]:
                worst = (f"{prefix}.{k}", e)
            assert e < 5 * TOL[precision], (prefix, k, e)
    print(f"[dropout/{precision}/B={B}] worst per-parameter gradient rel-err {worst[1]:.2e} ({worst[0]}); total norm {total:.4f}")
    # a second forward draws different masks
    m2 = step.plan.dropout_masks()
    assert not torch.equal(m2["fuse0"].cpu(), masks["fuse0"])


@pytest.mark.parametrize("B", [1, 4, 128, 1000])
def test_batch_sizes_and_eval_step(B):
    """fnd_eval_step at ragged and large batches (bf16): logits within tolerance of the oracle, argmax identical
    wherever the oracle's margin exceeds the tolerance."""
    f, c, fus, clf = build_pair(42, True, "bf16")
    f.eval(); c.eval()
    batch = O.make_batch(B, seed=5)
    step = FusedStep(f, c, B, precision="bf16", use_graph=(B == 128))
    step.load_batch(to_cuda(batch))
    step.eval_step()
    with torch.no_grad():
        ref = O.model_forward(fus, clf, batch)
    lg = step.logits().cpu()
    assert O.rel_err(lg, ref["logits"]) < TOL["bf16"]
    margin = (ref["logits"][:, 0] - ref["logits"][:, 1]).abs()
    sure = margin > 2 * TOL["bf16"] * ref["logits"].abs().max()
    assert torch.equal(lg.argmax(-1)[sure], ref["logits"].argmax(-1)[sure])
    row_loss = torch.nn.functional.cross_entropy(ref["logits"], batch["label"], reduction="none")
    assert O.rel_err(step.loss_rows().cpu(), row_loss) < TOL["bf16"]
    step.plan.check_error()


def test_missing_gnn_feat_raises_like_reference():
    f = CrossModalTransformer()
    x = {"text_features": torch.randn(2, 768), "audio_features": torch.randn(2, 128),
         "visual_features": torch.randn(2, 512), "temporal_features": torch.randn(2, 256), "gnn_feat": None}
    with pytest.raises(RuntimeError, match="shapes cannot be multiplied"):
        f(x)


def test_backward_after_overwritten_forward_raises():
    f, c, _, _ = build_pair(42, False, "bf16")
    f.train(); c.train()
    b1, b2 = to_cuda(O.make_batch(4, seed=1)), to_cuda(O.make_batch(4, seed=2))
    fo1, co1 = module_forward(f, c, b1)
    module_forward(f, c, b2)
    with pytest.raises(RuntimeError, match="overwritten"):
        torch.nn.functional.cross_entropy(co1["logits"], b1["label"]).backward()


@pytest.mark.parametrize("B", [200, 384])
def test_large_batch_train_step_matches_oracle(B):
    """Batches above 128 take the other code paths of the step (several M tiles, deep wgrad ring -> the one-CTA-per-SM
    GEMM variant with finalize as a kernel of its own): one fused step vs the oracle's step, fp32 mode, dropout off."""
    f, c, fus, clf = build_pair(42, True, "fp32", dropout_off=True)
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    batch = O.make_batch(B, seed=31)
    step = FusedStep(f, c, B, precision="fp32", use_graph=False)
    step.load_batch(to_cuda(batch))
    step.train_step()
    st = step.plan.state()
    step.plan.check_error()
    fo = {k: v.clone() for k, v in fus.items()}; co = {k: v.clone() for k, v in clf.items()}
    ref = O.train_step(fo, co, batch, O.AdamWState(), dropout=0.0)
    print(f"[B={B}] loss {st['loss']} vs {float(ref['loss'])}; norm {st['grad_norm']} vs {ref['grad_norm']}")
    assert abs(st["loss"] - float(ref["loss"])) / float(ref["loss"]) < TOL["fp32"]
    assert abs(st["grad_norm"] - ref["grad_norm"]) / ref["grad_norm"] < 2e-3
    assert O.rel_err(step.logits().cpu(), ref["logits"]) < TOL["fp32"]
    for name, mod, od in (("fuse_mlp.0.weight", f, fo), ("attn_tv.q.weight", f, fo), ("pre.0.weight", c, co), ("bypass.bias", c, co)):
        got = dict(mod.named_parameters())[name].detach().cpu()
        assert float((got - od[name]).abs().mean()) < 0.02 * 2e-4, name     # mean error << one lr-sized update


def test_eval_forward_b1024_fp32_and_bf16_vs_reference_oracle():
    """BASELINE.json configs[2]: inference-only eval forward, batch 1024, fp32 vs bf16 logits vs the reference
    (oracle pinned to it): rel-err <= 1e-3 / 2e-2 and identical argmax predictions."""
    B = 1024
    batch = O.make_batch(B, seed=9)
    ref = None
    for precision in ("fp32", "bf16"):
        f, c, fus, clf = build_pair(42, True, precision)
        f.eval(); c.eval()
        if ref is None:
            with torch.no_grad():
                ref = O.model_forward(fus, clf, batch)
        step = FusedStep(f, c, B, precision=precision, use_graph=True)
        step.load_batch(to_cuda(batch))
        step.eval_step()
        lg, pr = step.logits().cpu(), step.probs().cpu()
        step.plan.check_error()
        e = O.rel_err(lg, ref["logits"])
        margin = (ref["logits"][:, 0] - ref["logits"][:, 1]).abs()
        agree = (lg.argmax(-1) == ref["logits"].argmax(-1))
        print(f"[eval B=1024/{precision}] logits rel-err {e:.2e}; argmax agreement {int(agree.sum())}/{B}; "
              f"smallest reference margin among disagreements {float(margin[~agree].min()) if (~agree).any() else float('nan'):.2e}")
        assert e < TOL[precision]
        assert O.rel_err(pr, ref["probs"]) < TOL[precision]
        # predictions identical; a flip is only tolerated on a numerical tie (|margin| below the stated tolerance of the logits)
        assert bool(agree[margin > TOL[precision] * float(ref["logits"].abs().max())].all())
        if precision == "fp32":
            assert bool(agree.all())


def test_hidden_1024_forward_and_gradients_match_oracle(tmp_path):
    """hidden_dim is a YAML knob of the reference (fusion.yaml:2, classifier.yaml:3): the 1024-wide instantiations of
    the row kernels / GEMM tables against the (shape-generic) oracle, fp32 mode."""
    fy, cy = tmp_path / "fusion.yaml", tmp_path / "classifier.yaml"
    fy.write_text("hidden_dim: 1024\ndropout: 0.0\nuse_gnn: true\ngnn_dim: 128\n")
    cy.write_text("input_dim: 1024\nhidden_dim: 1024\ndropout: 0.0\nnum_classes: 2\nuse_aux: true\naux_dim: 2\n"
                  "node_trees: 6\nnode_depth: 4\nnode_tau: 10.0\ntemperature: 1.0\n")
    torch.manual_seed(5)
    f = CrossModalTransformer(config_path=str(fy), precision="fp32")
    c = DeepTruthClassifier(config_path=str(cy), precision="fp32")
    assert f.hidden == 1024 and f.fused_dim == 16 * 1024
    with torch.no_grad():
        g = torch.Generator().manual_seed(6)
        for n, p in c.named_parameters():
            if "gates" in n or "leaf_logits" in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g).to(p.device))
    for m in list(f.modules()) + list(c.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    B = 24
    batch = O.make_batch(B, seed=21)
    step = FusedStep(f, c, B, precision="fp32", use_graph=False)
    fus = {k: v.detach().cpu().clone() for k, v in f.state_dict().items()}
    clf = {k: v.detach().cpu().clone() for k, v in c.state_dict().items()}
    step.load_batch(to_cuda(batch))
    step.train_fwd_bwd()
    st = step.plan.state()
    step.plan.check_error()
    out, gf, gc = O.loss_and_grads(fus, clf, batch, dropout=0.0)
    print(f"[H=1024] loss {st['loss']} vs {float(out['loss'])}; logits rel-err {O.rel_err(step.logits().cpu(), out['logits']):.2e}")
    assert abs(st["loss"] - float(out["loss"])) / float(out["loss"]) < TOL["fp32"]
    assert O.rel_err(step.logits().cpu(), out["logits"]) < TOL["fp32"]
    assert O.rel_err(step.fused().cpu(), out["fused"]) < TOL["fp32"]
    eng = step.engine
    worst = 0.0
    for prefix, grads in (("fusion", gf), ("clf", gc)):
        for k, gref in grads.items():
            if float(gref.norm()) == 0:
                continue
            e = O.rel_err(eng.grad_view(f"{prefix}.{k}").cpu(), gref)
            worst = max(worst, e)
            assert e < 5 * TOL["fp32"], (prefix, k, e)
    print(f"[H=1024] worst per-parameter gradient rel-err {worst:.2e}")


def test_no_gnn_no_aux_configuration_matches_oracle(tmp_path):
    """fusion.yaml use_gnn: false (15 slots in fused_cat) and classifier.yaml use_aux: false (no rank-2 aux update,
    plain K = 512 pre.0): the other shape of the tables, against the oracle, fp32 mode."""
    fy, cy = tmp_path / "fusion.yaml", tmp_path / "classifier.yaml"
    fy.write_text("hidden_dim: 512\ndropout: 0.0\nuse_gnn: false\ngnn_dim: 128\n")
    cy.write_text("input_dim: 512\nhidden_dim: 512\ndropout: 0.0\nnum_classes: 2\nuse_aux: false\naux_dim: 2\n"
                  "node_trees: 6\nnode_depth: 4\nnode_tau: 10.0\ntemperature: 1.0\n")
    torch.manual_seed(8)
    f = CrossModalTransformer(config_path=str(fy), precision="fp32")
    c = DeepTruthClassifier(config_path=str(cy), precision="fp32")
    assert f.fused_dim == 15 * 512 and "gnn_proj.weight" not in f.state_dict()
    with torch.no_grad():
        g = torch.Generator().manual_seed(9)
        for n, p in c.named_parameters():
            if "gates" in n or "leaf_logits" in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g).to(p.device))
    for m in list(f.modules()) + list(c.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    B = 20
    batch = O.make_batch(B, seed=22)
    ob = dict(batch); ob["gnn_feat"] = None; ob["aux"] = None          # what the oracle (= the reference) sees
    step = FusedStep(f, c, B, precision="fp32", use_graph=False)
    fus = {k: v.detach().cpu().clone() for k, v in f.state_dict().items()}
    clf = {k: v.detach().cpu().clone() for k, v in c.state_dict().items()}
    step.load_batch(to_cuda(batch))
    step.train_fwd_bwd()
    st = step.plan.state()
    step.plan.check_error()
    fk, ck = O.trainable_keys()
    fl = {k: v.clone().requires_grad_(k in fk) for k, v in fus.items()}
    cl = {k: v.clone().requires_grad_(k in ck) for k, v in clf.items()}
    feats = {k: ob[k] for k in O.FEAT_KEYS}
    fo = O.fusion_forward(fl, feats, dropout=0.0)
    co = O.classifier_forward(cl, fo["fused"], None, dropout=0.0)
    loss = torch.nn.functional.cross_entropy(co["logits"], batch["label"])
    loss.backward()
    print(f"[no gnn/aux] loss {st['loss']} vs {loss.item()}; logits rel-err {O.rel_err(step.logits().cpu(), co['logits'].detach()):.2e}")
    assert abs(st["loss"] - loss.item()) / loss.item() < TOL["fp32"]
    assert O.rel_err(step.logits().cpu(), co["logits"].detach()) < TOL["fp32"]
    eng = step.engine
    for prefix, params in (("fusion", fl), ("clf", cl)):
        for k, p in params.items():
            if p.grad is None or float(p.grad.norm()) == 0:
                continue
            e = O.rel_err(eng.grad_view(f"{prefix}.{k}").cpu(), p.grad)
            assert e < 5 * TOL["fp32"], (prefix, k, e)


def _load_large(name):
    z = np.load(os.path.join(GOLD, "large", name + ".npz"))
    batch = O.make_batch(int(z["meta_batch"]), seed=int(z["meta_data_seed"]))
    return z, batch


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_b1024_against_reference_outputs(precision):
    """BASELINE.json configs[2] against outputs of the UNMODIFIED reference itself (tests/golden/large/eval_b1024.npz):
    logits rel-err <= 1e-3 (fp32 mode) / 2e-2 (bf16), identical argmax wherever the reference's margin is not a tie."""
    z, batch = _load_large("eval_b1024")
    f, c, _, _ = build_pair(int(z["meta_seed"]), True, precision)
    f.eval(); c.eval()
    B = int(z["meta_batch"])
    step = FusedStep(f, c, B, precision=precision, use_graph=True)
    step.load_batch(to_cuda(batch))
    step.eval_step()
    step.plan.check_error()
    ref_logits = torch.from_numpy(z["eval.logits"])
    lg = step.logits().cpu()
    assert O.rel_err(lg, ref_logits) < TOL[precision]
    assert O.rel_err(step.probs().cpu(), torch.from_numpy(z["eval.probs"])) < TOL[precision]
    assert O.rel_err(step.fused().cpu().double().sum(-1), torch.from_numpy(z["eval.fused_rowsum"])) < TOL[precision]
    rs = step.plan.buffer("rowstat", torch.float32, (B, 16)).cpu()
    for col, key in ((0, "semantic_conflict"), (1, "emotion_intensity"), (2, "temporal_delay")):
        assert O.rel_err(rs[:, col], torch.from_numpy(z["eval.forensic." + key])) < TOL[precision], key
    margin = (ref_logits[:, 0] - ref_logits[:, 1]).abs()
    agree = lg.argmax(-1) == ref_logits.argmax(-1)
    assert bool(agree[margin > TOL[precision] * float(ref_logits.abs().max())].all())
    if precision == "fp32":
        assert bool(agree.all())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_b128_against_reference_outputs(precision):
    """BASELINE.json configs[1] (the bench batch) against the reference's own loss and per-parameter gradient norms
    (tests/golden/large/train_b128.npz), dropout off."""
    z, batch = _load_large("train_b128")
    f, c, _, _ = build_pair(int(z["meta_seed"]), True, precision, dropout_off=True)
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    B = int(z["meta_batch"])
    step = FusedStep(f, c, B, precision=precision, use_graph=True)
    step.load_batch(to_cuda(batch))
    step.train_fwd_bwd()
    st = step.plan.state()
    step.plan.check_error()
    tol = TOL[precision]
    assert abs(st["loss"] - float(z["train.loss"])) / float(z["train.loss"]) < tol
    assert abs(st["grad_norm"] - float(z["train.grad_norm"])) / float(z["train.grad_norm"]) < max(tol, 2e-3)
    eng = step.engine
    worst = 0.0
    for key in z.files:
        if not key.startswith("gnorm."):
            continue
        name = key[len("gnorm."):]
        ref = float(z[key])
        if ref == 0.0:
            continue
        got = float(eng.grad_view(name).double().norm())
        worst = max(worst, abs(got - ref) / ref)
        assert abs(got - ref) / ref < (5e-4 if precision == "fp32" else 5e-2), (name, got, ref)
    print(f"[train_b128/{precision}] loss {st['loss']} vs {float(z['train.loss'])}; worst gradient-norm rel-err {worst:.2e}")


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_overlap_is_bit_identical(use_graph):
    """fnd_train_step_overlap (fuse_mlp weight gradients on a side stream under the rest of the backward chain) runs the same
    tiles and writes the same norm slots as fnd_train_step: parameters, Adam moments and losses must be BIT-identical after
    several steps with dropout on (forensic_trainer.py:285-298 is the step both replace)."""
    batch = O.make_batch(128, seed=21)
    res = []
    for early in (False, True):
        torch.manual_seed(3)
        f, c = CrossModalTransformer(precision="bf16"), DeepTruthClassifier(precision="bf16")
        f.train(); c.train()
        st = FusedStep(f, c, 128, precision="bf16", use_graph=use_graph)
        st.wg_early = early
        st.load_batch({k: v.cuda() for k, v in batch.items()})
        losses = []
        for _ in range(4):
            st.train_step()
            losses.append(st.plan.state()["loss"])
        st.plan.check_error()
        eng = st.engine
        res.append((losses, eng.params.clone(), eng.adam_m.clone(), eng.adam_v.clone()))
    (l0, p0, m0, v0), (l1, p1, m1, v1) = res
    assert l0 == l1
    assert torch.equal(p0, p1) and torch.equal(m0, m1) and torch.equal(v0, v1)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_cluster_splitk_exchange_is_bit_identical(precision):
    """Split-K partial tiles exchanged through distributed shared memory inside thread-block clusters (csrc/fnd_gemm.cuh) vs
    through the L2 workspace: same plan (same tiles and splits), same fixed summation order -> parameters, Adam moments and
    losses must be BIT-identical after several optimizer steps with dropout on."""
    from ultrafnd_git_b200 import _lib
    lib = _lib.load()
    batch = O.make_batch(128, seed=33)
    res = []
    try:
        for on in (1, 0):
            lib.fnd_debug_set_cluster_splitk(1)           # plans are laid out with the cluster-era tile / split choices
            torch.manual_seed(5)
            f, c = CrossModalTransformer(precision=precision), DeepTruthClassifier(precision=precision)
            f.train(); c.train()
            st = FusedStep(f, c, 128, precision=precision, use_graph=False)
            st.load_batch({k: v.cuda() for k, v in batch.items()})
            lib.fnd_debug_set_cluster_splitk(on)          # ... and launched through one exchange or the other
            losses = []
            for _ in range(3):
                st.train_step()
                losses.append(st.plan.state()["loss"])
            st.plan.check_error()
            eng = st.engine
            res.append((losses, eng.params.clone(), eng.adam_m.clone(), eng.adam_v.clone()))
    finally:
        lib.fnd_debug_set_cluster_splitk(1)
    (l1, p1, m1, v1), (l0, p0, m0, v0) = res
    assert l1 == l0
    assert torch.equal(p1, p0) and torch.equal(m1, m0) and torch.equal(v1, v0)
