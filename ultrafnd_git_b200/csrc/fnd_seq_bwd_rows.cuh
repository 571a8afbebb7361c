// fnd_seq_bwd_rows.cuh — HBM-bound row kernels of the sequence front-end's BACKWARD pass (Tier B): LayerNorm backward
// (input gradient + per-CTA partial sums of the affine gradients), masked mean-pool backward, column sums (bias
// gradients), the fixed-order reduction of those partial sums, and the attention backward's row prologue
// (D = rowsum(dO * O), logsumexp converted to exp2 units and padded). 128-bit global accesses throughout.
//
// No counterpart in the reference; checked against torch autograd over the self-oracle oracle/seq_oracle.py.
// The pool backward is the derivative of the one reference semantic, src/core_blocks/text_blocks.py:81-86.
#pragma once
#include "fnd_seq_rows.cuh"

namespace fnd {

// dt = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (t - mean) * rstd     (t = the LayerNorm input)
// dgamma = sum_rows dy * xhat, dbeta = sum_rows dy: every warp keeps its share in registers over the rows it walks, the
// CTA's eight warps meet in shared memory (fixed order) and the CTA writes ONE partial row pair part[cta][0|1][d]
// (seq_reduce_partials_kernel adds them in CTA order: deterministic, no atomics).
struct LnBwdParams {
  const __nv_bfloat16* t; int t_pitch;           // forward input of the LayerNorm (pre-normalisation), [M, t_pitch]
  const __nv_bfloat16* dy; int dy_pitch;         // upstream gradient [M, dy_pitch]
  const float* gamma;
  float eps;
  __nv_bfloat16* dt; int dt_pitch;               // gradient w.r.t. t
  float* part;                                   // [gridDim.x][2][d] partial dgamma / dbeta
  int M, d;
};
template <int kChunks>
__global__ void __launch_bounds__(256) seq_layernorm_bwd_kernel(const LnBwdParams P) {
  extern __shared__ float ln_red[];              // [8 warps][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = P.d >> 3;
  const float inv_d = 1.0f / static_cast<float>(P.d);
  float g[kChunks][8], dg[kChunks][8], db[kChunks][8];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int ch = c * 32 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[c][j] = 0.f; db[c][j] = 0.f; g[c][j] = 0.f; }
    if (ch < nchunk) {
      const float4 g0 = ldg_f4(P.gamma + ch * 8), g1 = ldg_f4(P.gamma + ch * 8 + 4);
      g[c][0] = g0.x; g[c][1] = g0.y; g[c][2] = g0.z; g[c][3] = g0.w; g[c][4] = g1.x; g[c][5] = g1.y; g[c][6] = g1.z; g[c][7] = g1.w;
    }
  }
  const int wstride = gridDim.x * 8;
#pragma unroll 1
  for (int row = blockIdx.x * 8 + warp; row < P.M; row += wstride) {
    const __nv_bfloat16* tr = P.t + static_cast<size_t>(row) * P.t_pitch;
    const __nv_bfloat16* dr = P.dy + static_cast<size_t>(row) * P.dy_pitch;
    float v[kChunks][8], u[kChunks][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = c * 32 + lane;
      if (ch < nchunk) {
        unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(tr + ch * 8)), v[c]);
        unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(dr + ch * 8)), u[c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[c][j];
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      if (c * 32 + lane < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[c][j] -= mean;
          sq = fmaf(v[c][j], v[c][j], sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + P.eps);
    float s1 = 0.f, s2 = 0.f;                     // sum g, sum g * xhat
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      if (c * 32 + lane < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = v[c][j] * rstd;
          const float gy = u[c][j] * g[c][j];
          dg[c][j] = fmaf(u[c][j], xh, dg[c][j]);
          db[c][j] += u[c][j];
          v[c][j] = xh;
          u[c][j] = gy;
          s1 += gy;
          s2 = fmaf(gy, xh, s2);
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
    __nv_bfloat16* or_ = P.dt + static_cast<size_t>(row) * P.dt_pitch;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = c * 32 + lane;
      if (ch < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (u[c][j] - s1 - v[c][j] * s2);
        *reinterpret_cast<uint4*>(or_ + ch * 8) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
  // ---- the CTA's partial dgamma / dbeta: warps -> shared memory -> one row pair per CTA ----
  float* mine = ln_red + static_cast<size_t>(warp) * 2 * P.d;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int ch = c * 32 + lane;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { mine[ch * 8 + j] = dg[c][j]; mine[P.d + ch * 8 + j] = db[c][j]; }
    }
  }
  __syncthreads();
  float* out = P.part + static_cast<size_t>(blockIdx.x) * 2 * P.d;
  for (int i = threadIdx.x; i < 2 * P.d; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += ln_red[static_cast<size_t>(w) * 2 * P.d + i];
    out[i] = s;
  }
}

// out[i] = sum_p part[p][i], fixed order (deterministic); n floats per partial row (multiple of 4), 128-bit loads.
// A CTA owns 128 consecutive floats: 32 column lanes x 8 part lanes (part lane y adds parts y, y + 8, ...), then the eight
// partial sums meet in shared memory in lane order. (One thread per column walking all parts serialised ~300 dependent
// loads on two CTAs: 57 us per call for the 2 x 1024 LayerNorm gradients.)
__global__ void __launch_bounds__(256) seq_reduce_partials_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out) {
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i4 = blockIdx.x * 32 + tx;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 * 4 < n) {
#pragma unroll 4
    for (int p = ty; p < nparts; p += 8) {
      const float4 v = ldcg_f4(part + static_cast<size_t>(p) * n + i4 * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i4 * 4 < n) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = red[w][tx];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i4 * 4) = s;
  }
}

// Column sums of a bf16 matrix (bias gradients): part[blockIdx.y][n] = sum over this CTA's row slice of x[m][n].
// Grid (cdiv(N, 256), row slices); 256 threads = 8 row lanes x 32 column lanes of 8 elements.
struct ColsumParams {
  const __nv_bfloat16* x; int x_pitch;
  int M, N;
  float* part;                                   // [gridDim.y][N]
};
__global__ void __launch_bounds__(256) seq_colsum_kernel(const ColsumParams P) {
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + tx * 8;
  const int rows_per = (P.M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(P.M, r0 + rows_per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < P.N) {
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += 8) {
      float v[8];
      unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(P.x + static_cast<size_t>(r) * P.x_pitch + col)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < P.N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    P.part[static_cast<size_t>(blockIdx.y) * P.N + c] = s;
  }
}

// Masked mean-pool backward: dx[b,l,:] = m[b,l] / clamp_min(sum_l m[b,l], 1e-6) * dpooled[b,:]   (text_blocks.py:81-86)
// One warp per token row; the sample's valid count is recomputed by the CTA (rows of a CTA may span two samples at most
// when L >= 8, so every warp counts its own sample: L bytes through L1).
struct PoolBwdParams {
  const float* dp; int dp_pitch;                 // [B, dp_pitch] gradient of the pooled vector (fp32)
  const unsigned char* mask;                     // [B, L] or null
  const int* len;                                // [B] or null
  int B, L, d;
  __nv_bfloat16* dx; int dx_pitch;               // [B*L, dx_pitch]
};
__global__ void __launch_bounds__(256) seq_masked_mean_pool_bwd_kernel(const PoolBwdParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = P.d >> 3;
  const long long M = static_cast<long long>(P.B) * P.L;
  int cached_b = -1;
  float inv_cnt = 0.f;
#pragma unroll 1
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < M; row += static_cast<long long>(gridDim.x) * 8) {
    const int b = static_cast<int>(row / P.L), l = static_cast<int>(row - static_cast<long long>(b) * P.L);
    const int Lv = P.len ? min(max(P.len[b], 0), P.L) : P.L;
    const unsigned char* mrow = P.mask ? P.mask + static_cast<size_t>(b) * P.L : nullptr;
    if (b != cached_b) {
      float cnt = 0.f;
      if (mrow) {
        for (int i = lane; i < Lv; i += 32) cnt += mrow[i] ? 1.f : 0.f;
        cnt = warp_sum(cnt);
      } else {
        cnt = static_cast<float>(Lv);
      }
      inv_cnt = 1.f / fmaxf(cnt, 1e-6f);
      cached_b = b;
    }
    const bool valid = l < Lv && (!mrow || mrow[l] != 0);
    const float w = valid ? inv_cnt : 0.f;
    __nv_bfloat16* xr = P.dx + static_cast<size_t>(row) * P.dx_pitch;
    const float* dr = P.dp + static_cast<size_t>(b) * P.dp_pitch;
    for (int ch = lane; ch < nchunk; ch += 32) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        const float4 a = ldg_f4(dr + ch * 8), c = ldg_f4(dr + ch * 8 + 4);
        o = make_uint4(pack_bf16x2(a.x * w, a.y * w), pack_bf16x2(a.z * w, a.w * w), pack_bf16x2(c.x * w, c.y * w), pack_bf16x2(c.z * w, c.w * w));
      }
      *reinterpret_cast<uint4*>(xr + ch * 8) = o;
    }
  }
}

// Attention backward prologue: per (sample, query row, head)
//   Dp[b,h,q]    = sum_d dO[b,q,h,d] * O[b,q,h,d]
//   lse2p[b,h,q] = lse[b,h,q] * log2(e), +inf when the row had no valid key (lse = -inf) or q >= Lq (padding) — a +inf
//                  offset makes every recomputed probability exactly 0 in the backward kernels.
// Both outputs are [B, H, Lqp] with Lqp a multiple of 64 (16-byte aligned 64-row slices for the bulk copies of the dK/dV
// kernel). One warp per (b, q) row: a head's 64 columns are 8 lanes x 8 elements, reduced with three shuffles.
struct AttnBwdPrepParams {
  const __nv_bfloat16* o; int o_pitch;           // forward output [B*Lq, o_pitch], head h at columns h*64
  const __nv_bfloat16* d_o; int do_pitch;        // its gradient
  const float* lse;                              // [B, H, Lq]
  int B, H, Lq, Lqp;
  float* Dp; float* lse2p;                       // [B, H, Lqp]
};
__global__ void __launch_bounds__(256) seq_attn_bwd_prep_kernel(const AttnBwdPrepParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(P.B) * P.Lqp;
#pragma unroll 1
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < rows; r += static_cast<long long>(gridDim.x) * 8) {
    const int b = static_cast<int>(r / P.Lqp), q = static_cast<int>(r - static_cast<long long>(b) * P.Lqp);
    if (q >= P.Lq) {
      for (int h = lane; h < P.H; h += 32) {
        const size_t o = (static_cast<size_t>(b) * P.H + h) * P.Lqp + q;
        P.Dp[o] = 0.f;
        P.lse2p[o] = INFINITY;
      }
      continue;
    }
    const __nv_bfloat16* orow = P.o + (static_cast<size_t>(b) * P.Lq + q) * P.o_pitch;
    const __nv_bfloat16* drow = P.d_o + (static_cast<size_t>(b) * P.Lq + q) * P.do_pitch;
    for (int h0 = 0; h0 < P.H; h0 += 4) {          // 32 lanes x 8 elements = 4 heads per pass
      const int h = h0 + (lane >> 3);
      float s = 0.f;
      if (h < P.H) {
        float a[8], c[8];
        unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(orow + h0 * 64 + lane * 8)), a);
        unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(drow + h0 * 64 + lane * 8)), c);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(a[j], c[j], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((lane & 7) == 0 && h < P.H) {
        const size_t o = (static_cast<size_t>(b) * P.H + h) * P.Lqp + q;
        const float l = P.lse[(static_cast<size_t>(b) * P.H + h) * P.Lq + q];
        P.Dp[o] = s;
        P.lse2p[o] = (l == -INFINITY) ? INFINITY : l * 1.44269504088896340736f;
      }
    }
  }
}

}  // namespace fnd

namespace fnd {

// ---------------------------------------------------------------------------------------------------------------
// Optimizer step of the sequence front-end over ONE flat fp32 parameter / gradient / moment range (the front-end's own
// mirror of Tier A's flat-arena AdamW, csrc/fnd_optim.cuh): global-norm clip + decoupled-weight-decay Adam in one pass,
// 128-bit accesses, 28 B of HBM traffic per parameter. Semantics = torch.nn.utils.clip_grad_norm_(max_norm) followed by
// torch.optim.AdamW (what the reference's trainer uses for its own parameters, src/training/forensic_trainer.py:173-177,
// 292-298); there is no reference optimizer for these parameters (the front-end does not exist there).
// ---------------------------------------------------------------------------------------------------------------
// part[blockIdx.x * 4] = sum of squares of this CTA's slice (three zero pads keep the 128-bit partial-sum reduction happy)
__global__ void __launch_bounds__(256) seq_sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldcg_f4(g + 4 * i);
    s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    *reinterpret_cast<float4*>(part + 4 * static_cast<size_t>(blockIdx.x)) = make_float4(t, 0.f, 0.f, 0.f);
  }
}

struct SeqAdamParams {
  float* w; const float* g; float* m; float* v;
  long long n4;                                  // number of float4 groups
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2, max_norm, grad_scale;
  const float* sumsq;                            // device scalar: sum of squares of g BEFORE grad_scale
};
__global__ void __launch_bounds__(256) seq_adamw_kernel(const SeqAdamParams P) {
  // clip coefficient of clip_grad_norm_: min(1, max_norm / (||g|| + 1e-6)), with ||g|| of the SCALED gradient
  const float norm = sqrtf(__ldg(P.sumsq)) * P.grad_scale;
  const float coef = (P.max_norm > 0.f) ? fminf(1.f, P.max_norm / (norm + 1e-6f)) : 1.f;
  const float gs = P.grad_scale * coef;
  const float decay = 1.f - P.lr * P.weight_decay;
  const float step = P.lr / P.bc1;
  const float inv_sqrt_bc2 = rsqrtf(P.bc2);
  const uint64_t pol = l2_policy_evict_first();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P.n4; i += stride) {
    const float4 g4 = ld_f4_policy(P.g + 4 * i, pol);
    float4 w4 = ld_f4_policy(P.w + 4 * i, pol), m4 = ld_f4_policy(P.m + 4 * i, pol), v4 = ld_f4_policy(P.v + 4 * i, pol);
    float* wp = &w4.x; float* mp = &m4.x; float* vp = &v4.x; const float* gp = &g4.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * gs;
      mp[j] = fmaf(P.beta1, mp[j], (1.f - P.beta1) * gj);
      vp[j] = fmaf(P.beta2, vp[j], (1.f - P.beta2) * gj * gj);
      const float denom = fmaf(sqrtf(vp[j]), inv_sqrt_bc2, P.eps);
      wp[j] = fmaf(-step, __fdividef(mp[j], denom), wp[j] * decay);
    }
    st_f4_policy(P.w + 4 * i, w4, pol);
    st_f4_policy(P.m + 4 * i, m4, pol);
    st_f4_policy(P.v + 4 * i, v4, pol);
  }
}

}  // namespace fnd
