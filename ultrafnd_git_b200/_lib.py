"""Loader and builder for libfnd_b200.so — the C-ABI library that holds every CUDA kernel of this package.

The library is built in-tree with plain ``nvcc`` for sm_100a only (no torch headers, no JIT cache) and is
bound with ``ctypes``: the Python side passes raw device pointers, sizes and the CUDA stream handle, exactly
what a cgo/JNI/ctypes binding in any other host language would pass.

There is deliberately no fallback: if the library is missing or a symbol declared in ``include/fnd_b200.h``
is absent, loading raises.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
import sys
from typing import Dict, List, Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
LIB_PATH = os.path.join(_PKG_DIR, "libfnd_b200.so")
HEADER_PATH = os.path.join(_REPO_DIR, "include", "fnd_b200.h")
SEQ_HEADER_PATH = os.path.join(_REPO_DIR, "include", "fnd_seq_b200.h")
HEADER_PATHS = [HEADER_PATH, SEQ_HEADER_PATH]
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _sources() -> List[str]:
    return [os.path.join(CSRC_DIR, "fnd_api.cu"), os.path.join(CSRC_DIR, "fnd_seq_api.cu")]


def _deps() -> List[str]:
    out = list(HEADER_PATHS)
    for fn in sorted(os.listdir(CSRC_DIR)):
        if fn.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC_DIR, fn))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libfnd_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["--threads", "2"] + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + _sources()
    proc = subprocess.run(cmd, cwd=_REPO_DIR, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


def declared_symbols() -> List[str]:
    """Every function name declared in include/fnd_b200.h and include/fnd_seq_b200.h."""
    names = set()
    for path in HEADER_PATHS:
        with open(path, "r", encoding="utf-8") as f:
            src = f.read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        names.update(re.findall(r"\b(fnd_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


_lib: Optional[ctypes.CDLL] = None

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_size_t = ctypes.c_size_t
_c_float = ctypes.c_float
_c_u64 = ctypes.c_uint64


class FndDims(ctypes.Structure):
    """Mirror of ``fnd_dims`` (include/fnd_b200.h)."""
    _fields_ = [("hidden", _c_int), ("d_text", _c_int), ("d_audio", _c_int), ("d_visual", _c_int),
                ("d_temporal", _c_int), ("d_gnn", _c_int), ("use_gnn", _c_int), ("aux_dim", _c_int),
                ("trees", _c_int), ("depth", _c_int), ("fusion_dropout", _c_float), ("clf_dropout", _c_float),
                ("tree_dropout", _c_float), ("node_tau", _c_float)]


class FndInputs(ctypes.Structure):
    """Mirror of ``fnd_inputs`` (include/fnd_b200.h)."""
    _fields_ = [("x", _c_void_p * 5), ("pitch", _c_int * 5), ("aux", _c_void_p), ("aux_pitch", _c_int),
                ("labels", _c_void_p), ("gather", _c_void_p)]


def _signatures() -> Dict[str, tuple]:
    P = ctypes.POINTER
    ll = ctypes.c_longlong
    return {
        "fnd_version": (_c_int, []),
        "fnd_build_arch": (ctypes.c_char_p, []),
        "fnd_param_count": (_c_int, [P(FndDims)]),
        "fnd_param_info": (_c_int, [P(FndDims), _c_int, ctypes.c_char_p, _c_int, P(ll), P(_c_int), P(_c_int),
                                    P(_c_int), P(_c_int)]),
        "fnd_arena_total_elems": (ll, [P(FndDims)]),
        "fnd_arena_hot_elems": (ll, [P(FndDims)]),
        "fnd_arena_shadow_elems": (ll, [P(FndDims)]),
        "fnd_arena_shadow_buffer_elems": (ll, [P(FndDims)]),
        "fnd_plan_create": (_c_int, [P(FndDims), _c_int, _c_int, P(_c_void_p)]),
        "fnd_plan_destroy": (None, [_c_void_p]),
        "fnd_plan_workspace_bytes": (_c_size_t, [_c_void_p]),
        "fnd_plan_bind": (_c_int, [_c_void_p] * 9),
        "fnd_plan_set_grad_mirror": (_c_int, [_c_void_p, _c_void_p]),
        "fnd_plan_buffer_offset": (ll, [_c_void_p, ctypes.c_char_p]),
        "fnd_plan_buffer_bytes": (ll, [_c_void_p, ctypes.c_char_p]),
        "fnd_set_hyper": (_c_int, [_c_void_p, _c_float, _c_float, _c_float, _c_float, _c_float, _c_float, _c_void_p]),
        "fnd_set_lr": (_c_int, [_c_void_p, _c_float, _c_void_p]),
        "fnd_set_seed": (_c_int, [_c_void_p, ctypes.c_ulonglong, _c_void_p]),
        "fnd_set_loss_scale": (_c_int, [_c_void_p, _c_float, _c_void_p]),
        "fnd_set_loss_mirror": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p]),
        "fnd_collect_rows": (_c_int, [_c_void_p, _c_int, ctypes.c_longlong, ctypes.c_longlong, _c_void_p, _c_void_p, _c_void_p,
                                      _c_void_p, _c_void_p, _c_void_p]),
        "fnd_refresh_shadows": (_c_int, [_c_void_p, _c_void_p]),
        "fnd_fusion_forward": (_c_int, [_c_void_p, P(FndInputs), _c_int, _c_void_p]),
        "fnd_classifier_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
        "fnd_ce_loss_fwd_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
        "fnd_classifier_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
        "fnd_fusion_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p]),
        "fnd_clip_adamw_step": (_c_int, [_c_void_p, _c_int, _c_void_p]),
        "fnd_train_fwd_bwd": (_c_int, [_c_void_p, P(FndInputs), _c_void_p]),
        "fnd_train_step": (_c_int, [_c_void_p, P(FndInputs), _c_void_p]),
        "fnd_train_step_overlap": (_c_int, [_c_void_p, P(FndInputs), _c_void_p, _c_void_p]),
        "fnd_eval_step": (_c_int, [_c_void_p, P(FndInputs), _c_void_p]),
        "fnd_launch_count": (_c_int, [_c_void_p, ctypes.c_char_p]),
        "fnd_debug_set_launch_limit": (_c_int, [_c_void_p, _c_int]),
        "fnd_debug_set_cluster_splitk": (_c_int, [_c_int]),
        "fnd_dp_bind": (_c_int, [_c_void_p, _c_int, _c_int, P(ctypes.c_ulonglong), ll, ll, ll, ll, ll, ll, _c_int, ctypes.c_ulonglong, ll,
                                 _c_void_p, ll, _c_void_p, ll]),
        "fnd_dp_stage_bytes": (ll, [_c_void_p, _c_int, _c_int]),
        "fnd_dp_shard_ranges": (_c_int, [_c_void_p, _c_int, _c_int, P(ll), P(ll)]),
        "fnd_train_step_dp": (_c_int, [_c_void_p, P(FndInputs), _c_void_p, _c_void_p, _c_int]),
        "fnd_dp_flush": (_c_int, [_c_void_p, _c_void_p]),
        "fnd_dp_optimizer_step": (_c_int, [_c_void_p, _c_void_p]),
        "fnd_profile_begin": (_c_int, [_c_void_p, _c_void_p]),
        "fnd_profile_end": (_c_int, [_c_void_p, _c_void_p, ctypes.c_char_p, P(_c_float), _c_int, P(_c_int)]),
        "fnd_export_dropout_mask": (_c_int, [_c_void_p, _c_int, _c_void_p, ll, _c_void_p]),
        "fnd_check_error": (_c_int, [_c_void_p, _c_void_p]),
        # ---- sequence front-end (include/fnd_seq_b200.h)
        "fnd_seq_init": (_c_int, []),
        "fnd_seq_pair_clusters": (_c_int, []),
        "fnd_seq_debug_attn_stamps": (_c_int, [_c_void_p]),
        "fnd_seq_debug_gemm_stamps": (_c_int, [_c_void_p]),
        "fnd_seq_cast_bf16": (_c_int, [_c_void_p, _c_void_p, ll, _c_void_p]),
        "fnd_seq_linear": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int,
                                    _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
        "fnd_seq_layernorm": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_float, _c_void_p, _c_int, _c_int,
                                       _c_int, _c_void_p]),
        "fnd_seq_coattn_forward": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int,
                                            _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_int,
                                            _c_void_p, _c_void_p, _c_void_p]),
        "fnd_seq_masked_mean_pool": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int,
                                              _c_void_p, _c_int, _c_void_p]),
        "fnd_seq_coattn_backward_workspace": (_c_size_t, [_c_int] * 3),
        "fnd_seq_coattn_backward": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int,
                                             _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                             _c_int, _c_int, _c_int, _c_int, _c_float,
                                             _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int,
                                             _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
        "fnd_seq_layernorm_backward_workspace": (_c_size_t, [_c_int] * 2),
        "fnd_seq_layernorm_backward": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_float, _c_void_p, _c_int,
                                                _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
        "fnd_seq_colsum_workspace": (_c_size_t, [_c_int] * 2),
        "fnd_seq_colsum": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
        "fnd_seq_masked_mean_pool_backward": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                                                       _c_void_p, _c_int, _c_void_p]),
        "fnd_seq_grad_sumsq_workspace": (_c_size_t, [ll]),
        "fnd_seq_grad_sumsq": (_c_int, [_c_void_p, ll, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
        "fnd_seq_adamw_step": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, ll, _c_float, _c_float, _c_float, _c_float,
                                        _c_float, _c_int, _c_float, _c_float, _c_void_p, _c_void_p]),
        "fnd_seq_wgrad_workspace": (_c_size_t, [_c_int] * 3),
        "fnd_seq_wgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int,
                                   _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
        "fnd_gemm_bf16_async": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int,
                                         _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p, _c_void_p]),
        "fnd_gemm_scratch_bytes": (_c_size_t, [_c_int] * 4),
        "fnd_gemm_bf16_probe": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int,
                                         _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                         _c_void_p, _c_size_t, _c_void_p, _c_void_p, _c_int]),
        "fnd_gemm_bf16": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_int, _c_int,
                                   _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_void_p, _c_size_t, _c_void_p]),
    }


def load() -> ctypes.CDLL:
    """Load the library (building it first if the sources are newer) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        if os.environ.get("FND_NO_BUILD"):
            raise RuntimeError(f"{LIB_PATH} is missing or stale and FND_NO_BUILD is set")
        build()
    lib = ctypes.CDLL(LIB_PATH)
    sigs = _signatures()
    for name in declared_symbols():
        if not hasattr(lib, name):
            raise RuntimeError(f"libfnd_b200.so does not export {name} declared in include/*.h")
        if name not in sigs:
            raise RuntimeError(f"{name} is declared in include/*.h but has no ctypes signature in _lib.py")
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = sigs[name]
    _lib = lib
    return lib


class FndError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status == 0:
        return
    if status <= -1000:
        raise FndError(f"{what}: CUDA runtime error {-status - 1000}")
    if status < 0:
        raise FndError(f"{what}: invalid argument / host error {status}")
    raise FndError(f"{what}: device-side error code {status}")
