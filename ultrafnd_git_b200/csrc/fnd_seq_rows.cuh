// fnd_seq_rows.cuh — HBM-bound row kernels of the sequence front-end (Tier B): fp32 -> bf16 cast, LayerNorm and masked
// mean-pool. All global accesses are 128-bit (uint4 / float4), rows are contiguous and 16-byte aligned.
//
// Masked mean-pool follows the reference's only use of it, src/core_blocks/text_blocks.py:81-86:
//   sum_l x[l] * m[l] / clamp_min(sum_l m[l], 1e-6).  LayerNorm has no counterpart in the reference.
#pragma once
#include "fnd_common.cuh"

namespace fnd {

// y[i] = bf16(x[i]); n is a multiple of 8. Each thread converts 8 elements: two 128-bit loads, one 128-bit store.
__global__ void __launch_bounds__(256) seq_cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n8) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = ldcg_f4(x + 8 * i), b = ldcg_f4(x + 8 * i + 4);
    *reinterpret_cast<uint4*>(y + 8 * i) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
    v[2 * t] = __low2float(h2);
    v[2 * t + 1] = __high2float(h2);
  }
}

// LayerNorm over the last dimension (d <= 2048, d % 8 == 0), optionally of a SUM: y = LN(x + r).
//   y = (t - mean) * rsqrt(var + eps) * gamma + beta,  t = x (+ r)     (biased variance, fp32 statistics, two-pass in registers)
// A warp owns whole rows (the row lives in registers) and WALKS the row list with gamma / beta held in registers:
// re-reading the affine parameters per row (8 KB through L1 per 2 KB row) made L1, not HBM, the limit (ncu: l1tex 72 %,
// 3.1 TB/s). All global accesses are 128-bit.
constexpr int kLnMaxChunks = 8;                  // 8 chunks x 32 lanes x 8 elements = 2048
struct LnParams {
  const __nv_bfloat16* x; int x_pitch;
  const __nv_bfloat16* r; int r_pitch;           // optional residual (null: plain LayerNorm)
  const float* gamma; const float* beta;
  float eps;
  __nv_bfloat16* y; int y_pitch;
  int M, d;
};
template <int kChunks>                           // 8-element chunks per lane: d <= kChunks * 256
__global__ void __launch_bounds__(256) seq_layernorm_kernel(const LnParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = P.d >> 3;                   // 8-element chunks in the row
  const float inv_d = 1.0f / static_cast<float>(P.d);
  // gamma / beta live in SHARED memory (16 KB): per-row re-reads through L1 made L1 the limit (ncu: l1tex 72 %), holding
  // them in registers (64 per thread at d = 1024) left two CTAs per SM and 32 KB of rows in flight per SM — 3.2 TB/s.
  // From shared memory the kernel needs ~64 registers, four CTAs per SM keep 64 KB in flight.
  __shared__ float4 sg[kLnMaxChunks * 64], sb[kLnMaxChunks * 64];
  for (int i = threadIdx.x; i < (P.d >> 2); i += blockDim.x) {
    sg[i] = ldg_f4(P.gamma + 4 * i);
    sb[i] = ldg_f4(P.beta + 4 * i);
  }
  __syncthreads();
  const int wstride = gridDim.x * 8;
#pragma unroll 1
  for (int row = blockIdx.x * 8 + warp; row < P.M; row += wstride) {
    const __nv_bfloat16* xr = P.x + static_cast<size_t>(row) * P.x_pitch;
    const __nv_bfloat16* rr = P.r ? P.r + static_cast<size_t>(row) * P.r_pitch : nullptr;
    float v[kChunks][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = c * 32 + lane;
      if (ch < nchunk) unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(xr + ch * 8)), v[c]);
    }
    if (rr) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int ch = c * 32 + lane;
        if (ch < nchunk) {
          float t[8];
          unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(rr + ch * 8)), t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[c][j] += t[j];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      if (c * 32 + lane < nchunk) {
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
          const float t = v[c][j] - mean;
          sq = fmaf(t, t, sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + P.eps);
    __nv_bfloat16* yr = P.y + static_cast<size_t>(row) * P.y_pitch;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = c * 32 + lane;
      if (ch < nchunk) {
        const float4 g0 = sg[2 * ch], g1 = sg[2 * ch + 1], b0 = sb[2 * ch], b1 = sb[2 * ch + 1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[c][j] - mean) * rstd, gg[j], bb[j]);
        *reinterpret_cast<uint4*>(yr + ch * 8) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
}

// Masked mean over the sequence axis: out[b, :] = sum_l x[b,l,:] m[b,l] / clamp_min(sum_l m[b,l], 1e-6).
// Grid (d / 64, B); 256 threads = 32 row lanes x 8 column lanes; a row lane reads 128 contiguous bytes (64 bf16) per row
// and strides over the sequence, partial sums meet in shared memory in a fixed order (deterministic).
struct PoolParams {
  const __nv_bfloat16* x; int x_pitch;           // [B*L, x_pitch]
  const unsigned char* mask;                     // [B, L] or null (all valid)
  const int* len;                                // [B] prefix length or null
  int B, L, d;
  float* out_f32; int f32_pitch;                 // [B, f32_pitch] or null
  __nv_bfloat16* out_bf; int bf_pitch;           // [B, bf_pitch] or null
};
__global__ void __launch_bounds__(256) seq_masked_mean_pool_kernel(const PoolParams P) {
  __shared__ float part[32][65];
  __shared__ float cnt_s[32];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int col = c0 + tx * 8;
  const int L = P.len ? min(max(P.len[b], 0), P.L) : P.L;
  const unsigned char* mrow = P.mask ? P.mask + static_cast<size_t>(b) * P.L : nullptr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float cnt = 0.f;
  const bool col_ok = col < P.d;
#pragma unroll 4
  for (int l = ty; l < L; l += 32) {
    const float w = mrow ? (mrow[l] ? 1.f : 0.f) : 1.f;
    cnt += w;
    if (col_ok && w != 0.f) {
      float v[8];
      unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(P.x + (static_cast<size_t>(b) * P.L + l) * P.x_pitch + col)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[ty][tx * 8 + j] = acc[j];
  if (tx == 0) cnt_s[ty] = cnt;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f, n = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { s += part[r][threadIdx.x]; n += cnt_s[r]; }
    const int c = c0 + threadIdx.x;
    if (c < P.d) {
      const float o = s / fmaxf(n, 1e-6f);
      if (P.out_f32) P.out_f32[static_cast<size_t>(b) * P.f32_pitch + c] = o;
      if (P.out_bf) P.out_bf[static_cast<size_t>(b) * P.bf_pitch + c] = __float2bfloat16_rn(o);
    }
  }
}

}  // namespace fnd
