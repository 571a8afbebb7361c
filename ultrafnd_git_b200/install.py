"""Drop-in installation: make the reference's own import paths resolve to the B200 implementation.

    import ultrafnd_git_b200.install as fnd; fnd.install()
    # now, unchanged reference code such as run_train_eval.py / scripts/smoke_test_v2.py does
    from src.models.fusion.cross_modal_transformer import CrossModalTransformer     # -> B200 module
    from src.models.fusion.deep_truth_classifier import DeepTruthClassifier          # -> B200 module
    from src.training.forensic_trainer import TrainConfig, ForensicTrainer           # -> B200 trainer

Only the three hot-path modules are replaced; everything else under ``src`` (data pipeline, feature
extractors, GNN utilities) keeps coming from the reference checkout when it is on ``sys.path``. When it is not, empty
parent packages are created so the three names above still import.
"""
from __future__ import annotations

import importlib
import sys
import types

HOT_MODULES = {
    "src.models.fusion.cross_modal_transformer": ("ultrafnd_git_b200.modules", ["CrossModalTransformer", "ForensicCoAttention"]),
    "src.models.fusion.deep_truth_classifier": ("ultrafnd_git_b200.modules", ["DeepTruthClassifier", "NODEEnsemble", "_ObliviousTree"]),
    "src.training.forensic_trainer": ("ultrafnd_git_b200.trainer", ["TrainConfig", "ForensicTrainer", "SimpleGCN",
                                                                    "CachedTensorDataset", "build_adj_from_ocr"]),
}


def _ensure_package(name: str) -> None:
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
    except Exception:
        pkg = types.ModuleType(name)
        pkg.__path__ = []  # type: ignore[attr-defined]
        sys.modules[name] = pkg
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, pkg)


def install() -> None:
    for full, (src_mod, names) in HOT_MODULES.items():
        parts = full.split(".")
        for i in range(1, len(parts)):
            _ensure_package(".".join(parts[:i]))
        impl = importlib.import_module(src_mod)
        shim = types.ModuleType(full)
        shim.__doc__ = f"B200 drop-in for {full} (provided by {src_mod})"
        for n in names:
            setattr(shim, n, getattr(impl, n))
        sys.modules[full] = shim
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], shim)


def uninstall() -> None:
    for full in HOT_MODULES:
        sys.modules.pop(full, None)
