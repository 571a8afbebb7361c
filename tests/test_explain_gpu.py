"""SURVEY.md §8 f4 (second half): DeepTruthClassifier.feature_importance — Gradient x Input through the library's own
backward — against autograd over the oracle's restated classifier (deep_truth_classifier.py:189-211)."""
import pytest
import torch

from oracle import fnd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_feature_importance_matches_oracle_autograd(precision, tol):
    from ultrafnd_git_b200.modules import DeepTruthClassifier
    _, clf_p = O.init_params(42)
    O.perturb_node_head(clf_p)
    B = 24
    g = torch.Generator().manual_seed(3)
    fused = torch.randn(B, 512, generator=g)
    aux = torch.rand(B, 2, generator=g)
    clf = DeepTruthClassifier(precision=precision)
    clf.load_state_dict(clf_p)
    clf.eval()
    imp, agg = clf.feature_importance(fused.cuda(), aux.cuda(), class_idx=1, aggregate=True)
    assert imp.shape == (B, 514) and agg.shape == (514,)
    # oracle: autograd over the restated classifier (eval mode: no dropout), same Gradient x Input definition
    f, a = fused.clone().requires_grad_(True), aux.clone().requires_grad_(True)
    out = O.classifier_forward({k: v.clone() for k, v in clf_p.items()}, f, a, dropout=0.0)
    out["logits"][:, 1].sum().backward()
    x = torch.cat([f, a], -1).detach()
    ref = (torch.cat([f.grad, a.grad], -1) * x).abs()
    e_f = O.rel_err(imp[:, :512].cpu(), ref[:, :512])
    e_a = O.rel_err(imp[:, 512:].cpu(), ref[:, 512:])
    print(f"feature_importance[{precision}]: fused part rel-err {e_f:.2e}, aux part rel-err {e_a:.2e}")
    assert e_f < tol and e_a < tol
    assert O.rel_err(agg.cpu(), ref.mean(0)) < tol
    # class 0 and the no-aggregate form
    imp0, none = clf.feature_importance(fused.cuda(), aux.cuda(), class_idx=0, aggregate=False)
    assert none is None and imp0.shape == (B, 514) and not torch.allclose(imp0, imp)
