"""CPU tests of the host-side logic: C-ABI surface, arena layout, drop-in module API, metrics, graph helper,
install shim. No kernel is launched here (there is no GPU in the build container)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import fnd_oracle as O
from ultrafnd_git_b200 import _lib, engine as E
from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier, pair_modules

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _lib.declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.fnd_build_arch() == b"sm_100a"
    assert lib.fnd_version() >= 100


def test_arena_layout_matches_state_dict_shapes():
    tab = E.param_table(E.Dims())
    names = {p.name: p for p in tab}
    for prefix, shapes in (("fusion.", O.fusion_param_shapes()), ("clf.", O.classifier_param_shapes())):
        for k, shp in shapes.items():
            assert tuple(names[prefix + k].shape) == tuple(shp), k
    assert len(tab) == len(O.fusion_param_shapes()) + len(O.classifier_param_shapes())
    # no overlaps; hot tensors form a prefix [0, n_hot)
    spans = sorted((p.offset, p.offset + p.numel, p.hot) for p in tab)
    for (a0, a1, _), (b0, _, _) in zip(spans, spans[1:]):
        assert a1 <= b0
    eng = E.Engine(E.Dims(), device=torch.device("cpu"))
    assert all((p.offset + p.numel <= eng.n_hot) == p.hot for p in tab)
    fk, ck = O.trainable_keys()
    assert {p.name for p in tab if p.hot} == {"fusion." + k for k in fk} | {"clf." + k for k in ck}
    assert sum(p.numel for p in tab if p.hot) == 12745949          # parameters that receive gradients
    assert eng.n_hot % 64 == 0 and eng.n_shadow % 64 == 0


def test_unsupported_dims_fail_loudly():
    with pytest.raises(NotImplementedError):
        E.Engine(E.Dims(hidden=384), device=torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        E.Engine(E.Dims(aux_dim=3), device=torch.device("cpu"))


def test_modules_mirror_reference_api_and_refuse_cpu_compute():
    torch.manual_seed(42)
    f, c = CrossModalTransformer(), DeepTruthClassifier()
    assert list(f.state_dict().keys()) == list(O.fusion_param_shapes().keys())
    assert list(c.state_dict().keys()) == list(O.classifier_param_shapes().keys())
    assert (f.hidden, f.dropout, f.use_gnn, f.gnn_dim, f.fused_dim) == (512, 0.1, True, 128, 8192)
    assert (c.hidden, c.aux_dim, c.node_trees, c.node_depth, c.node_tau) == (512, 2, 6, 4, 10.0)
    assert not c.node.trees[0].tau.requires_grad and c.temperature.requires_grad
    batch = O.make_batch(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        f({k: batch[k] for k in O.FEAT_KEYS})
    with pytest.raises(RuntimeError, match="CUDA"):
        c(torch.randn(2, 512), batch["aux"])
    # load_state_dict writes through to the shared arena after pairing
    eng = pair_modules(f, c)
    fus, clf = O.init_params(3)
    f.load_state_dict(fus); c.load_state_dict(clf)
    assert torch.equal(eng.view("fusion.fuse_mlp.0.weight"), fus["fuse_mlp.0.weight"])
    assert torch.equal(eng.view("clf.bypass.weight"), clf["bypass.weight"])
    assert f.text_proj.weight.data_ptr() == eng.view("fusion.text_proj.weight").data_ptr()
    v0 = eng.param_version()
    with torch.no_grad():
        f.text_proj.bias.add_(1.0)
    assert eng.param_version() != v0          # in-place torch updates invalidate the bf16 shadows


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_same_seed_gives_reference_identical_initial_weights():
    """The drop-in consumes the global torch RNG in the reference's construction order."""
    code = f"""
import sys, os, torch, hashlib
sys.path.insert(0, {REF!r}); os.chdir({REF!r}); os.environ["HF_HUB_OFFLINE"] = "1"
torch.manual_seed(123)
from src.models.fusion.cross_modal_transformer import CrossModalTransformer
from src.models.fusion.deep_truth_classifier import DeepTruthClassifier
f = CrossModalTransformer("configs/model_configs/fusion.yaml"); c = DeepTruthClassifier("configs/model_configs/classifier.yaml")
for pre, m in (("f", f), ("c", c)):
    for k, v in m.state_dict().items():
        print(pre, k, hashlib.md5(v.detach().contiguous().numpy().tobytes()).hexdigest())
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         env={**os.environ, "PYTHONDONTWRITEBYTECODE": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    ref = {tuple(l.split()[:2]): l.split()[2] for l in out.stdout.splitlines() if l[:2] in ("f ", "c ")}
    import hashlib
    torch.manual_seed(123)
    f, c = CrossModalTransformer(), DeepTruthClassifier()
    assert len(ref) == len(f.state_dict()) + len(c.state_dict())
    for pre, m in (("f", f), ("c", c)):
        for k, v in m.state_dict().items():
            assert hashlib.md5(v.detach().contiguous().numpy().tobytes()).hexdigest() == ref[(pre, k)], (pre, k)


def test_metrics_match_sklearn():
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score, roc_auc_score
    from ultrafnd_git_b200.metrics import aggregate_epoch_metrics
    g = np.random.RandomState(0)
    y = g.randint(0, 2, 500)
    p = np.round(g.rand(500), 2)                   # rounding creates ties
    sc, td, ei = g.rand(500), g.rand(500), g.rand(500)
    m = aggregate_epoch_metrics(y, p, {"semantic_conflict": sc, "temporal_delay": td, "emotion_intensity": ei})
    pred = (p >= 0.5).astype(int)
    assert m["accuracy"] == pytest.approx(accuracy_score(y, pred))
    assert m["auc"] == pytest.approx(roc_auc_score(y, p), abs=1e-12)
    assert m["precision"] == pytest.approx(precision_score(y, pred))
    assert m["recall"] == pytest.approx(recall_score(y, pred))
    assert m["f1"] == pytest.approx(f1_score(y, pred))
    assert m["cmcs"] == pytest.approx(1 - np.clip(0.5 * (sc + td), 0, 1).mean())
    assert m["dfdr"] == pytest.approx(recall_score(y, pred))
    assert aggregate_epoch_metrics(np.ones(5, int), np.full(5, 0.9))["auc"] == 0.5      # single class -> chance level
    assert aggregate_epoch_metrics(np.array([], int), np.array([]))["accuracy"] == 0.0   # empty split


def test_ocr_graph_matches_pairwise_jaccard():
    from ultrafnd_git_b200.trainer import build_adj_from_ocr
    g = np.random.RandomState(1)
    vocab = [f"t{i}" for i in range(12)]
    sets = [set(g.choice(vocab, size=g.randint(0, 5), replace=False).tolist()) for _ in range(40)]
    a = build_adj_from_ocr(sets, 0.12)
    for i in range(40):
        for j in range(40):
            if i == j:
                exp = 1.0
            elif not sets[i] and not sets[j]:
                exp = 0.0
            else:
                exp = float(len(sets[i] & sets[j]) / (len(sets[i] | sets[j]) + 1e-9) >= 0.12)
            assert a[i, j] == exp, (i, j)


def test_install_shim_routes_reference_import_paths():
    from ultrafnd_git_b200 import install as inst, modules, trainer
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    try:
        inst.install()
        from src.models.fusion.cross_modal_transformer import CrossModalTransformer as A
        from src.models.fusion.deep_truth_classifier import DeepTruthClassifier as B
        from src.training.forensic_trainer import ForensicTrainer as C, TrainConfig as D
        assert A is modules.CrossModalTransformer and B is modules.DeepTruthClassifier
        assert C is trainer.ForensicTrainer and D is trainer.TrainConfig
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_train_config_fields_match_reference():
    from ultrafnd_git_b200.trainer import TrainConfig, synthetic_cache
    cfg = TrainConfig(data_root="x", ocr_phrase_pkl=None)
    assert (cfg.batch_size, cfg.epochs, cfg.lr, cfg.weight_decay, cfg.grad_clip, cfg.early_stop_patience) == (16, 8, 2e-4, 1e-4, 5.0, 3)
    c = synthetic_cache(64)
    assert c["text"].shape == (64, 768) and c["aux"].shape == (64, 2) and len(c["ocr_sets"]) == 64
    assert sum(len(s) for s in c["split"]) == 64


def test_bench_reference_arm_emits_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--batch", "16"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "cpu_baseline", "e2e", "config"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_dp_shard_ranges_partition_the_hot_arena():
    """fnd_dp_shard_ranges (host-only): for every world size the per-rank ranges are disjoint, 4-element aligned and cover
    [0, hot) exactly; segment 0 of every rank lies inside fuse_mlp.0.weight (the early-reduced range)."""
    import ctypes
    from ultrafnd_git_b200 import _lib, engine as E
    lib = _lib.load()
    dims = E.Dims()
    cd = dims.to_c()
    h = ctypes.c_void_p()
    assert lib.fnd_plan_create(ctypes.byref(cd), 8, 0, ctypes.byref(h)) == 0
    n_hot = lib.fnd_arena_hot_elems(ctypes.byref(cd))
    table = {p.name: p for p in E.param_table(dims)}
    f0 = table["fusion.fuse_mlp.0.weight"]
    try:
        for world in range(1, 9):
            covered = np.zeros(n_hot, dtype=np.int8)
            for rank in range(world):
                lo, hi = (ctypes.c_longlong * 3)(), (ctypes.c_longlong * 3)()
                assert lib.fnd_dp_shard_ranges(h, rank, world, lo, hi) == 3
                for s in range(3):
                    assert 0 <= lo[s] <= hi[s] <= n_hot and lo[s] % 4 == 0 and hi[s] % 4 == 0
                    covered[lo[s]:hi[s]] += 1
                assert f0.offset <= lo[0] and hi[0] <= table["clf.pre.0.weight"].offset
            assert (covered == 1).all(), f"world {world}: ranges must tile the hot arena exactly once"
            assert lib.fnd_dp_stage_bytes(h, world, 0) >= 4 * n_hot and lib.fnd_dp_stage_bytes(h, world, 1) * 2 == lib.fnd_dp_stage_bytes(h, world, 0)
    finally:
        lib.fnd_plan_destroy(h)


def test_sequence_frontend_host_logic_and_no_cpu_fallback():
    """Tier B host side (no GPU): parameter names equal the self-oracle's, the gradient production order covers every
    parameter exactly once with LayerNorm (weight, bias) pairs adjacent (the backward reduces them as one row pair), and
    both the module and its trainer refuse to compute on the CPU (no fallback)."""
    from oracle import seq_oracle as SO
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd, SequenceTrainer
    fe = SequenceFrontEnd(128, 2)
    names = [k for k, _ in fe.named_parameters()]
    shapes = SO.param_shapes(SO.FAKESV_STREAMS, SO.FAKESV_BLOCKS, 128)
    assert sorted(names) == sorted(shapes) and all(tuple(dict(fe.named_parameters())[k].shape) == shapes[k] for k in names)
    order = [k for bucket in fe.grad_order() for k in bucket]
    assert sorted(order) == sorted(names) and len(set(order)) == len(order)
    assert len(fe.grad_order()) == 2 + len(fe.block_pairs)                # heads | one bucket per block | embeddings
    for i, k in enumerate(order):
        if (".ln." in k or k.startswith("embed_ln.")) and k.endswith(".weight"):
            assert order[i + 1] == k[:-len("weight")] + "bias", (k, order[i + 1])
    assert all(dict(fe.named_parameters())[k].numel() % 4 == 0 for k in names)      # 16-byte aligned flat views
    batch = SO.make_batch(SO.FAKESV_STREAMS, {"text": 8, "frames": 8, "audio": 8, "c3d": 8}, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        fe(batch)
    with pytest.raises(RuntimeError, match="CUDA"):
        SequenceTrainer(fe)
    with pytest.raises(NotImplementedError):
        SequenceFrontEnd(100, 2)                                           # head dimension must be 64


def test_shipped_library_contains_blackwell_native_sass():
    """The built libfnd_b200.so really is sm_100a tensor-core code (runs without a GPU: cuobjdump only): tcgen05.mma in the
    latency GEMM, the cta_group::2 forms in the CTA-pair projection GEMM, cluster barriers for the distributed-shared-memory
    split-K exchange, TMA loads, and no legacy mma.sync (HMMA) anywhere."""
    import re
    import shutil
    import subprocess
    from ultrafnd_git_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    _lib.build(force=False)
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = []
        elif cur is not None and "/*" in line:
            per[cur].append(line)

    def count(kernel_substr, pattern):
        return sum(len(re.findall(pattern, "\n".join(v))) for k, v in per.items() if kernel_substr in k)

    assert count("fnd_gemm_kernel", r"\bUTCHMMA\b") > 0 and count("fnd_gemm_kernel", r"\bUTMALDG") > 0
    assert count("fnd_gemm_kernelILi0", r"\bUCGABAR_") > 0                     # cluster split-K (general variant only)
    assert count("seq_gemm2_kernel", r"\bUTCHMMA\.2CTA") > 0
    assert count("seq_gemm2_kernel", r"\bUTCBAR\.2CTA\.MULTICAST") > 0
    assert count("seq_gemm2_kernel", r"\bUTMALDG\.2D\.2CTA") > 0
    assert count("seq_attn_fwd_kernel", r"\bUTCHMMA\b") > 0 and count("seq_attn_fwd_kernel", r"\bMUFU\.EX2") > 0
    assert count("", r"\bHMMA\b") == 0
