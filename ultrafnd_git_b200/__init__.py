"""ultrafnd_git_b200 — B200-native (sm_100a) fusion hot path for Ultrafnd.

Drop-in replacements for the reference's ``CrossModalTransformer`` / ``DeepTruthClassifier`` /
``ForensicTrainer`` backed by hand-written tcgen05/TMA CUDA kernels behind a C ABI (``include/fnd_b200.h``).
"""
__version__ = "0.1.0"
