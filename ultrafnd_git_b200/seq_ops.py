"""Thin ctypes wrappers over the sequence front-end entry points of libfnd_b200.so (include/fnd_seq_b200.h).

torch supplies device memory and the current stream only; every operation below is one call into the C ABI and raises
when CUDA or the library is unavailable (no eager fallback). bf16 matrices are passed as 2-D (or flattened) row-major
tensors whose row stride is the pitch.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import check


def _need_cuda(*ts: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("ultrafnd_git_b200 sequence ops run on CUDA (sm_100a) only: no CPU fallback exists")
        dev = t.device
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _as2d(t: torch.Tensor) -> torch.Tensor:
    if t.dim() == 3:
        assert t.stride(0) == t.shape[1] * t.stride(1), "batch stride must equal rows * pitch"
        return t.as_strided((t.shape[0] * t.shape[1], t.shape[2]), (t.stride(1), t.stride(2)))
    return t


def new_err_flag(dev: torch.device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=dev)


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 (fnd_seq_cast_bf16). x contiguous, numel % 8 == 0."""
    dev = _need_cuda(x)
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
    check(_lib.load().fnd_seq_cast_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream(dev)), "fnd_seq_cast_bf16")
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
           act: int = 0, out: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
           want_bf16: bool = True, err: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = a @ w.T (+ bias) (+ resid) (-> GELU). a [M,K] bf16 (row stride = pitch), w [N,K] bf16, bias fp32 [N]."""
    dev = _need_cuda(a, w)
    a, w = _as2d(a), _as2d(w)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None and want_bf16:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    r2 = _as2d(resid) if resid is not None else None
    check(_lib.load().fnd_seq_linear(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                                     _ptr(r2), r2.stride(0) if r2 is not None else 0, int(act),
                                     _ptr(out), _as2d(out).stride(0) if out is not None else 0,
                                     _ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                     M, N, K, _ptr(err), _stream(dev)), "fnd_seq_linear")
    return out if out is not None else out_f32


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm(x + resid) over the last dimension (resid optional)."""
    dev = _need_cuda(x)
    x2 = _as2d(x)
    M, d = x2.shape
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
    o2 = _as2d(out)
    r2 = _as2d(resid) if resid is not None else None
    check(_lib.load().fnd_seq_layernorm(x2.data_ptr(), x2.stride(0), _ptr(r2), r2.stride(0) if r2 is not None else 0,
                                        gamma.data_ptr(), beta.data_ptr(), float(eps), o2.data_ptr(), o2.stride(0), M, d,
                                        _stream(dev)), "fnd_seq_layernorm")
    return out


def coattn_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, heads: int, Lq: int, Lk: int,
                   q_col0: int = 0, k_col0: int = 0, v_col0: int = 0, kv_len: Optional[torch.Tensor] = None,
                   kv_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                   lse: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                   err: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(Q K^T * scale + mask) V per head (d_k = 64). q / k / v: 2-D bf16 [B*L, pitch]; head h at col0 + 64 h."""
    dev = _need_cuda(q, k, v)
    q, k, v = _as2d(q), _as2d(k), _as2d(v)
    if out is None:
        out = torch.empty(B * Lq, heads * 64, dtype=torch.bfloat16, device=dev)
    if scale is None:
        scale = 1.0 / math.sqrt(64.0)
    if kv_mask is not None:
        assert kv_mask.dtype == torch.uint8 and kv_mask.is_contiguous() and tuple(kv_mask.shape) == (B, Lk)
    if kv_len is not None:
        assert kv_len.dtype == torch.int32 and kv_len.numel() == B
    o2 = _as2d(out)
    check(_lib.load().fnd_seq_coattn_forward(q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0,
                                             v.data_ptr(), v.stride(0), v_col0, _ptr(kv_len), _ptr(kv_mask), B, heads,
                                             Lq, Lk, float(scale), o2.data_ptr(), o2.stride(0), _ptr(lse), _ptr(err),
                                             _stream(dev)), "fnd_seq_coattn_forward")
    return out


def masked_mean_pool(x: torch.Tensor, B: int, L: int, mask: Optional[torch.Tensor] = None,
                     length: Optional[torch.Tensor] = None, want_bf16: bool = False):
    """sum(x * m) / clamp_min(sum(m), 1e-6) over the sequence axis (text_blocks.py:81-86). Returns fp32 [B, d]
    (and a bf16 copy when want_bf16)."""
    dev = _need_cuda(x)
    x2 = _as2d(x)
    d = x2.shape[1]
    out = torch.empty(B, d, dtype=torch.float32, device=dev)
    obf = torch.empty(B, d, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    check(_lib.load().fnd_seq_masked_mean_pool(x2.data_ptr(), x2.stride(0), _ptr(mask), _ptr(length), B, L, d,
                                               out.data_ptr(), d, _ptr(obf), d, _stream(dev)), "fnd_seq_masked_mean_pool")
    return (out, obf) if want_bf16 else out


# ----------------------------------------------------------------------------------------------------- backward pass
def _workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


def coattn_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor,
                    lse: torch.Tensor, B: int, heads: int, Lq: int, Lk: int, dq: torch.Tensor, dk: torch.Tensor,
                    dv: torch.Tensor, q_col0: int = 0, k_col0: int = 0, v_col0: int = 0, dq_col0: int = 0,
                    dk_col0: int = 0, dv_col0: int = 0, kv_len: Optional[torch.Tensor] = None,
                    kv_mask: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                    err: Optional[torch.Tensor] = None) -> None:
    """Fused attention backward (fnd_seq_coattn_backward): fills dq / dk / dv (bf16, head h at col0 + 64 h) from the
    forward's q / k / v / o / lse and the output gradient d_o."""
    dev = _need_cuda(q, k, v, o, d_o, lse, dq, dk, dv)
    q, k, v, o, d_o, dq, dk, dv = (_as2d(t) for t in (q, k, v, o, d_o, dq, dk, dv))
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * heads * Lq
    if scale is None:
        scale = 1.0 / math.sqrt(64.0)
    lib = _lib.load()
    ws = _workspace(lib.fnd_seq_coattn_backward_workspace(B, heads, Lq), dev)
    check(lib.fnd_seq_coattn_backward(q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
                                      v.stride(0), v_col0, o.data_ptr(), o.stride(0), d_o.data_ptr(), d_o.stride(0),
                                      lse.data_ptr(), _ptr(kv_len), _ptr(kv_mask), B, heads, Lq, Lk, float(scale),
                                      dq.data_ptr(), dq.stride(0), dq_col0, dk.data_ptr(), dk.stride(0), dk_col0,
                                      dv.data_ptr(), dv.stride(0), dv_col0, ws.data_ptr(), ws.numel(), _ptr(err),
                                      _stream(dev)), "fnd_seq_coattn_backward")


def layernorm_backward(t: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5,
                       dgb: Optional[torch.Tensor] = None):
    """Returns (dt bf16 [M, d], dgamma fp32 [d], dbeta fp32 [d]) for y = LayerNorm(t). ``dgb``: optional contiguous fp32
    [2, d] destination for (dgamma, dbeta) (a slice of a flat gradient buffer)."""
    dev = _need_cuda(t, dy, gamma)
    t2, dy2 = _as2d(t), _as2d(dy)
    M, d = t2.shape
    dt = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    if dgb is None:
        dgb = torch.empty(2, d, dtype=torch.float32, device=dev)
    assert dgb.is_contiguous() and dgb.dtype == torch.float32 and dgb.numel() == 2 * d
    lib = _lib.load()
    ws = _workspace(lib.fnd_seq_layernorm_backward_workspace(M, d), dev)
    check(lib.fnd_seq_layernorm_backward(t2.data_ptr(), t2.stride(0), dy2.data_ptr(), dy2.stride(0), gamma.data_ptr(),
                                         float(eps), dt.data_ptr(), d, dgb[0].data_ptr(), dgb[1].data_ptr(), M, d,
                                         ws.data_ptr(), ws.numel(), _stream(dev)), "fnd_seq_layernorm_backward")
    return dt, dgb[0], dgb[1]


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 column sums of a bf16 matrix (bias gradients)."""
    dev = _need_cuda(x)
    x2 = _as2d(x)
    M, N = x2.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=dev)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == N
    lib = _lib.load()
    ws = _workspace(lib.fnd_seq_colsum_workspace(M, N), dev)
    check(lib.fnd_seq_colsum(x2.data_ptr(), x2.stride(0), M, N, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
          "fnd_seq_colsum")
    return out


def masked_mean_pool_backward(dpooled: torch.Tensor, B: int, L: int, mask: Optional[torch.Tensor] = None,
                              length: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 [B*L, d] gradient of the token states from the fp32 [B, d] gradient of the pooled vectors."""
    dev = _need_cuda(dpooled)
    dpooled = dpooled.contiguous().float()
    d = dpooled.shape[1]
    dx = torch.empty(B * L, d, dtype=torch.bfloat16, device=dev)
    check(_lib.load().fnd_seq_masked_mean_pool_backward(dpooled.data_ptr(), d, _ptr(mask), _ptr(length), B, L, d,
                                                        dx.data_ptr(), d, _stream(dev)), "fnd_seq_masked_mean_pool_backward")
    return dx


def wgrad(dy: torch.Tensor, x: torch.Tensor, err: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW[N, K] (fp32) = dy[M, N]^T @ x[M, K] on the persistent tcgen05 GEMM (fnd_seq_wgrad): both operands are read in place
    with the token dimension as the reduction — MN-major descriptors — so nothing is transposed in memory."""
    dev = _need_cuda(dy, x)
    dy, x = _as2d(dy), _as2d(x)
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dy.stride(1) == 1 and x.stride(1) == 1
    if N % 8 or K % 8:
        raise NotImplementedError("wgrad needs N and K to be multiples of 8 (16-byte row pitches)")
    lib = _lib.load()
    if out is None:
        out = torch.empty(N, K, dtype=torch.float32, device=dev)
    assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (N, K)
    ws = _workspace(lib.fnd_seq_wgrad_workspace(M, N, K), dev)
    check(lib.fnd_seq_wgrad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, N, K, out.data_ptr(), K,
                            ws.data_ptr(), ws.numel(), _ptr(err), _stream(dev)), "fnd_seq_wgrad")
    return out


def grad_sumsq(g: torch.Tensor, out4: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out4[0] = sum of squares of the flat fp32 gradient range (fixed-order reduction); stays on the device."""
    dev = _need_cuda(g)
    assert g.is_contiguous() and g.dtype == torch.float32 and g.numel() % 4 == 0
    if out4 is None:
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.fnd_seq_grad_sumsq_workspace(g.numel()), dev)
    check(lib.fnd_seq_grad_sumsq(g.data_ptr(), g.numel(), out4.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)), "fnd_seq_grad_sumsq")
    return out4


def adamw_step(w: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, sumsq: torch.Tensor, step: int, lr: float,
               betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 0.0, grad_scale: float = 1.0) -> None:
    """clip_grad_norm_(max_norm) + torch.optim.AdamW semantics over one flat fp32 range, in place (fnd_seq_adamw_step)."""
    dev = _need_cuda(w, g, m, v, sumsq)
    n = w.numel()
    assert all(t.is_contiguous() and t.dtype == torch.float32 and t.numel() == n for t in (w, g, m, v)) and n % 4 == 0
    check(_lib.load().fnd_seq_adamw_step(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, float(lr), float(betas[0]),
                                         float(betas[1]), float(eps), float(weight_decay), int(step), float(max_norm),
                                         float(grad_scale), sumsq.data_ptr(), _stream(dev)), "fnd_seq_adamw_step")
