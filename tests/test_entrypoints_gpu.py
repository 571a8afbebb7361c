"""The reference's entry points on the B200 path through ``install.install()`` (missing item 7).

/root/reference does not travel to the GPU box, so the statements of run_train_eval.py:68-109 (``main()``) and of
scripts/smoke_test_v2.py:32-84 are replayed verbatim in structure — the same imports by the reference's module paths, the
same constructor arguments, the same calls and result keys — with the reference's data pipeline (out of scope, SURVEY.md
§2 #9) stubbed by a module that serves a synthetic feature cache."""
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def installed(tmp_path):
    import ultrafnd_git_b200.install as fnd
    from ultrafnd_git_b200.trainer import synthetic_cache
    fnd.install()
    # stub of src.data_pipeline.fakesv_dataset (FakeSVRawDataset / build_gnn_cache_from_raw_dataset): serves a cache dict
    # with exactly the keys the reference's builder returns (fakesv_dataset.py:242-252)
    stub = types.ModuleType("src.data_pipeline.fakesv_dataset")

    class FakeSVRawDataset:                                     # noqa: D401
        def __init__(self, root):
            self.root = root

    def build_gnn_cache_from_raw_dataset(raw, ocr_phrase_pkl=None, text_dim=768, audio_dim=128, visual_dim=512,
                                         temporal_dim=256, seed=42):
        assert (text_dim, audio_dim, visual_dim, temporal_dim) == (768, 128, 512, 256)
        return synthetic_cache(n=400, seed=seed)
    stub.FakeSVRawDataset, stub.build_gnn_cache_from_raw_dataset = FakeSVRawDataset, build_gnn_cache_from_raw_dataset
    saved = {k: sys.modules.get(k) for k in ("src.data_pipeline", "src.data_pipeline.fakesv_dataset")}
    pkg = types.ModuleType("src.data_pipeline")
    pkg.__path__ = []
    sys.modules["src.data_pipeline"] = pkg
    sys.modules["src.data_pipeline.fakesv_dataset"] = stub
    yield tmp_path
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    fnd.uninstall()


def test_smoke_test_v2_calls(installed):
    """scripts/smoke_test_v2.py:32-84 — model initialisation + forward shapes, trainer construction + test() keys."""
    from src.models.fusion.cross_modal_transformer import CrossModalTransformer
    from src.models.fusion.deep_truth_classifier import DeepTruthClassifier
    from src.training.forensic_trainer import ForensicTrainer, TrainConfig
    fusion = CrossModalTransformer("configs/model_configs/fusion.yaml")
    clf = DeepTruthClassifier("configs/model_configs/classifier.yaml")
    B = 2
    T, A = torch.randn(B, 768), torch.randn(B, 128)
    V, U = torch.randn(B, 512), torch.randn(B, 256)
    G = torch.randn(B, 128)
    fused_out = fusion({"text_features": T, "audio_features": A, "visual_features": V, "temporal_features": U, "gnn_feat": G})
    assert tuple(fused_out["fused"].shape) == (B, 512)
    res = clf(fused_out["fused"], torch.rand(B, 2))
    assert tuple(res["probs"].shape) == (B, 2)
    cfg = TrainConfig(data_root="/Volumes/SR_disk/FakeSV", ocr_phrase_pkl=None, out_dir=str(installed / "outputs_smoke"),
                      batch_size=4, epochs=0, lr=2e-4, weight_decay=1e-4, gnn_dim=128, gnn_overlap_thresh=0.12, seed=42,
                      use_mps=False, use_gnn=True, save_best=False)
    trainer = ForensicTrainer(cfg)
    res = trainer.test()
    for k in ("test_loss", "test_acc", "test_auc"):
        assert k in res


def test_run_train_eval_main_body(installed, capsys):
    """run_train_eval.py:68-109 — TrainConfig from the CLI defaults, ForensicTrainer(cfg), fit(), test(), result keys."""
    from src.training.forensic_trainer import ForensicTrainer, TrainConfig
    torch.manual_seed(42)
    cfg = TrainConfig(data_root="FakeSV", ocr_phrase_pkl=None, out_dir=str(installed / "outputs"), batch_size=16, epochs=3,
                      lr=2e-4, weight_decay=1e-4, gnn_dim=128, gnn_overlap_thresh=0.12, seed=42, use_mps=False, use_gnn=True,
                      save_best=True)
    trainer = ForensicTrainer(cfg)
    best = trainer.fit()
    results = trainer.test()
    out = capsys.readouterr().out
    assert "[Epoch 01] train_loss=" in out and "val_loss=" in out and "[Test] loss=" in out        # the reference's log lines
    assert set(results) >= {"test_loss", "test_acc", "test_auc", "test_precision", "test_recall", "test_f1", "test_cmcs", "test_dfdr"}
    assert 0.0 <= results["test_acc"] <= 1.0 and best > 0.5
    assert (installed / "outputs" / "best.pt").exists()
