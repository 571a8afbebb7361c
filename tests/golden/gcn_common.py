"""Deterministic inputs of the GCN-stage fixtures (shared by tests/golden/make_golden.py, which runs the REFERENCE on
them in the build container, and tests/test_gcn_*.py, which run on boxes without the reference checkout)."""
import numpy as np


def gcn_inputs(n, seed):
    """L2-normalised node features (N, 416) as forensic_trainer.py:192-194 builds them, and OCR phrase sets drawn from a
    Zipf-like vocabulary so that the Jaccard graph has isolated posts, small cliques and a few hubs."""
    g = np.random.RandomState(seed)
    X = g.randn(n, 416).astype(np.float32)
    X /= (np.linalg.norm(X, axis=1, keepdims=True) + 1e-9)
    vocab = 4 * n
    ocr = []
    for _ in range(n):
        k = g.randint(0, 5)
        ocr.append(set((g.zipf(1.3, size=k) % vocab).tolist()))
    return X, ocr
