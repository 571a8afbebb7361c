// fnd_dp.cuh — data-parallel optimizer step over NVLink peer memory (one process per GPU, batch-sharded replicas).
//
// The reference trains on one device (src/training/forensic_trainer.py:285-298); its clip_grad_norm_ + AdamW pair is
// what this file replaces when the batch is sharded over N GPUs. Instead of "NCCL all-reduce of the 51 MB fp32 gradient,
// then N identical AdamW passes over 357 MB each", the gradient exchange, the clip and the optimizer are ONE sharded
// sequence over symmetric (peer-mapped) buffers:
//
//   (copy engines)     the fuse_mlp.0 weight gradient (65 % of all gradient bytes) is complete two thirds of the way
//                      through the backward pass: each rank PUSHES the pieces its peers own into their staging
//                      buffers with cudaMemcpyAsync on a side stream (DMA engines over NVLink, no SM is taken from
//                      the backward kernels) and then raises a flag (dp_signal_kernel);
//   dp_reduce_kernel   rank r sums ITS 1/N slice of the gradient arena: the pushed pieces out of local staging, the
//                      rest straight out of every peer's memory (128-bit P2P loads over NVLink/NVSwitch — a
//                      reduce-scatter without a staging copy), keeps the reduced slice locally and publishes the
//                      slice's sum of squares to every peer;
//   dp_adamw_kernel    every rank adds the N partial sums in rank order (bit-identical clip coefficient everywhere),
//                      runs AdamW on its slice only (1/N of the 357 MB optimizer stream; fp32 master, m and v stay
//                      sharded, ZeRO-1 style) and writes the refreshed bf16 operand shadows — the only copy of the
//                      weights the forward/backward kernels read — plus the small fp32 parameters (biases, gates, ...)
//                      into EVERY rank's buffers with P2P stores (the all-gather);
//   dp_wait_kernel     blocks the stream until every peer's shadow writes have landed here.
//
// Cross-GPU ordering uses epoch flags in each rank's symmetric comm pad (st.release.sys / ld.acquire.sys): a flag
// holds the number of the last step for which the event happened, so nothing is ever reset and a captured CUDA graph
// can be replayed. All three kernels are ordinary stream-ordered launches; they only ever wait for events that peers
// produce without needing anything further from this rank, so the sequence cannot deadlock as long as every rank
// runs the same steps.
#pragma once
#include "fnd_optim.cuh"

namespace fnd {

constexpr int kDpMaxWorld = 8;
constexpr int kDpMaxSeg = 3;
// comm pad layout (uint32 words): [0,8) early-gradients-ready epochs | [8,16) all-gradients-ready | [16,24) partial
// norms ready | [24,32) shadows written | [32,40) float partial sums of squares | [40] local epoch counter |
// [41] local CTA counter | [42] float: this rank's partial of the early segment
constexpr int kPadReadyEarly = 0, kPadReadyLate = 8, kPadPartialReady = 16, kPadDone = 24, kPadPartial = 32, kPadEpoch = 40,
              kPadCounter = 41, kPadPartialEarly = 42;
constexpr int kPadWords = 64;

struct DpParams {
  int rank, world;
  const float* grads[kDpMaxWorld];        // gradient arena of every rank (peer-mapped)
  float* params[kDpMaxWorld];             // fp32 parameter arena of every rank
  __nv_bfloat16* sh_hi[kDpMaxWorld];      // bf16 operand shadows of every rank
  __nv_bfloat16* sh_lo[kDpMaxWorld];      // residual planes (fp32x3 mode) or null
  unsigned int* pad[kDpMaxWorld];         // comm pad of every rank
  float* stage[kDpMaxWorld];              // staging buffer of every rank: [world][piece_cap] pushed segment-0 pieces
  size_t piece_cap;                       // elements per staging slot
  float* gred;                            // local: reduced gradient slice [shard_hi - shard_lo]
  float* slots;                           // local: per-CTA sums of squares of dp_reduce_kernel
  // This rank's slice of [0, n_hot): one piece of each of (up to) three arena ranges — segment 0 is its share of the
  // "early" range (fuse_mlp.0.weight), segments 1 and 2 its shares of the ranges before and after it. gred holds the
  // reduced pieces back to back (seg_goff).
  int nseg;
  size_t seg_lo[kDpMaxSeg], seg_hi[kDpMaxSeg], seg_goff[kDpMaxSeg];
  AdamWParams a;                          // local p / m / v / state, shadow geometry
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// Spin until pad[base + p] >= epoch for every rank p (threads p < world of the calling warp), bounded like mbar_wait.
__device__ __forceinline__ void dp_wait_all(const unsigned int* pad, int base, int world, unsigned int epoch, int* err) {
  if (threadIdx.x < static_cast<unsigned>(world)) {
    long long t0 = 0;
    for (unsigned int spins = 1;; ++spins) {
      // epochs only grow; the signed difference tolerates wrap-around
      if (static_cast<int>(ld_acquire_sys(pad + base + threadIdx.x) - epoch) >= 0) break;
      if ((spins & 63u) == 0u) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ll) {       // ~10 s: a peer died or ran a different sequence
          if (err) atomicExch(err, 201);
          break;
        }
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// 1. reduce-scatter out of peer memory + slice norm
// ---------------------------------------------------------------------------------------------------------------
// One segment: W = padded world size, U = independent float4 per thread and iteration.
// `staged`: the peers' contributions were pushed into this rank's staging buffer (segment 0 of an overlapped step).
template <int W, int U>
__device__ __forceinline__ float dp_reduce_segment(const DpParams& d, int sg, bool staged) {
  const size_t n4 = (d.seg_hi[sg] - d.seg_lo[sg]) >> 2;
  const float* src[W];
#pragma unroll
  for (int p = 0; p < W; ++p) {
    src[p] = nullptr;
    if (p < d.world)
      src[p] = (staged && p != d.rank) ? d.stage[d.rank] + static_cast<size_t>(p) * d.piece_cap : d.grads[p] + d.seg_lo[sg];
  }
  float* out = d.gred + d.seg_goff[sg];
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  float ss = 0.f;
  for (size_t i4 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4; i4 += stride * U) {
    float4 t[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t j4 = i4 + u * stride;
#pragma unroll
      for (int p = 0; p < W; ++p)
        if (p < d.world && j4 < n4) t[u][p] = ld_peer_f4(src[p] + j4 * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t j4 = i4 + u * stride;
      if (j4 < n4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < W; ++p)      // summed in rank order
          if (p < d.world) { acc.x += t[u][p].x; acc.y += t[u][p].y; acc.z += t[u][p].z; acc.w += t[u][p].w; }
        *reinterpret_cast<float4*>(out + j4 * 4) = acc;
        ss += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
      }
    }
  }
  return ss;
}

// Raised on the side stream once this rank's pushes have been issued AND completed (stream order after the copies).
__global__ void __launch_bounds__(32) dp_signal_kernel(DpParams d, int bank) {
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  if (threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + bank + d.rank, epoch);
  }
}

// All segments of this rank's slice. `staged` != 0: segment 0 comes out of local staging (its pieces were pushed by the
// peers during the backward pass; wait for their kPadReadyEarly flags too).
__global__ void __launch_bounds__(256) dp_reduce_kernel(DpParams d, int staged) {
  __shared__ float red[8];
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  // all of this rank's gradients are complete (stream order): tell every peer
  if (blockIdx.x == 0 && threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + kPadReadyLate + d.rank, epoch);
  }
  if (staged) dp_wait_all(mypad, kPadReadyEarly, d.world, epoch, &d.a.state->err);
  dp_wait_all(mypad, kPadReadyLate, d.world, epoch, &d.a.state->err);

  float ss = 0.f;
  for (int sg = 0; sg < d.nseg; ++sg) {
    const bool st = staged && sg == 0;
    // eight 128-bit loads in flight per thread whatever the world size (NVLink latency x bandwidth needs MBs in flight)
    if (d.world <= 2) ss += dp_reduce_segment<2, 4>(d, sg, st);
    else if (d.world <= 4) ss += dp_reduce_segment<4, 2>(d, sg, st);
    else ss += dp_reduce_segment<8, 1>(d, sg, st);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    d.slots[blockIdx.x] = tot;
    __threadfence();
    is_last = (atomicAdd(mypad + kPadCounter, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  // last CTA: this rank's sum of squares (fixed order) -> every peer
  if (threadIdx.x < 32) {
    __threadfence();
    const double part = warp_reduce_slots(d.slots, static_cast<int>(gridDim.x));
    if (threadIdx.x == 0) mypad[kPadCounter] = 0u;
    if (threadIdx.x < static_cast<unsigned>(d.world)) {
      reinterpret_cast<float*>(d.pad[threadIdx.x])[kPadPartial + d.rank] = static_cast<float>(part);
      __threadfence_system();
      st_release_sys(d.pad[threadIdx.x] + kPadPartialReady + d.rank, epoch);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. clip + AdamW on the slice, shadows / small parameters written to every rank
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dp_publish4(const DpParams& d, size_t i, const float4& x) {
  const AdamWParams& a = d.a;
  if (i < a.n_shadow) {
    // bf16 operand copy (hi [, lo]) of a GEMM weight -> all ranks
    const uint32_t h0 = pack_bf16x2(x.x, x.y), h1 = pack_bf16x2(x.z, x.w);
    uint32_t l0 = 0, l1 = 0;
    if (a.sh_lo) {
      const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&h0), b2 = *reinterpret_cast<const __nv_bfloat162*>(&h1);
      l0 = pack_bf16x2(x.x - __low2float(a2), x.y - __high2float(a2));
      l1 = pack_bf16x2(x.z - __low2float(b2), x.w - __high2float(b2));
    }
#pragma unroll
    for (int p = 0; p < kDpMaxWorld; ++p) {
      if (p < d.world) {
        *reinterpret_cast<uint2*>(d.sh_hi[p] + i) = make_uint2(h0, h1);
        if (a.sh_lo) *reinterpret_cast<uint2*>(d.sh_lo[p] + i) = make_uint2(l0, l1);
      }
    }
  } else {
    // small fp32 parameters (biases, gates, thresholds, leaf tables, evidence MLPs) are read as fp32 by the kernels
#pragma unroll
    for (int p = 0; p < kDpMaxWorld; ++p)
      if (p < d.world && p != d.rank) *reinterpret_cast<float4*>(d.params[p] + i) = x;
  }
  if (i + 4 > a.rp_begin && i < a.rp_end) {
    // re-pitched pre.0.weight shadow (+ its fp32 aux columns, which the epilogue reads from the arena)
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const size_t rp_off = static_cast<size_t>(a.rp_hi - a.sh_hi);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t e = i + q;
      if (e >= a.rp_begin && e < a.rp_end) {
        const size_t r = (e - a.rp_begin) / a.rp_cols, c = (e - a.rp_begin) % a.rp_cols;
        __nv_bfloat16 h, l;
        split_bf16(xs[q], h, l);
        for (int p = 0; p < d.world; ++p) {
          d.sh_hi[p][rp_off + r * a.rp_pitch + c] = h;
          if (a.sh_lo) d.sh_lo[p][rp_off + r * a.rp_pitch + c] = l;
          if (c >= static_cast<size_t>(a.rp_cols - 2) && p != d.rank) d.params[p][e] = xs[q];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) dp_adamw_kernel(DpParams d) {
  __shared__ float s_norm;
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  DevState* S = d.a.state;
  dp_wait_all(mypad, kPadPartialReady, d.world, epoch, &S->err);
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int p = 0; p < d.world; ++p) tot += static_cast<double>(reinterpret_cast<volatile float*>(mypad)[kPadPartial + p]);
    s_norm = static_cast<float>(sqrt(tot));
  }
  __syncthreads();
  const float norm = s_norm;
  const float coef = clip_coef_of(S->max_norm, norm);
  const float lr = S->lr, b1 = S->beta1, b2 = S->beta2, eps = S->eps;
  const int t = S->step + 1;
  const float bc1 = 1.0f - powf(b1, static_cast<float>(t)), bc2 = 1.0f - powf(b2, static_cast<float>(t));
  const float decay = 1.0f - lr * S->weight_decay;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const AdamWParams& a = d.a;
  const uint64_t pol = l2_policy_evict_first();
  for (int sg = 0; sg < d.nseg; ++sg) {
  const size_t n4 = (d.seg_hi[sg] - d.seg_lo[sg]) >> 2;
  const float* gsrc = d.gred + d.seg_goff[sg];
  for (size_t i4 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4;
       i4 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t i = d.seg_lo[sg] + i4 * 4;
    float4 p = ld_f4_policy(a.p + i, pol);
    const float4 g4 = *reinterpret_cast<const float4*>(gsrc + i4 * 4);
    float4 m = ld_f4_policy(a.m + i, pol);
    float4 v = ld_f4_policy(a.v + i, pol);
    float* pp = &p.x; float* mp = &m.x; float* vp = &v.x; const float* gp = &g4.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g = gp[q] * coef;
      pp[q] *= decay;
      mp[q] = b1 * mp[q] + (1.0f - b1) * g;
      vp[q] = b2 * vp[q] + (1.0f - b2) * g * g;
      const float denom = sqrtf(vp[q]) * inv_sqrt_bc2 + eps;
      pp[q] -= step_size * (mp[q] / denom);
    }
    st_f4_policy(a.p + i, p, pol);
    st_f4_policy(a.m + i, m, pol);
    st_f4_policy(a.v + i, v, pol);
    dp_publish4(d, i, p);
  }
  }
  // every P2P store of this CTA is ordered before its counter bump; the last CTA tells the peers
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(mypad + kPadCounter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) {
    S->grad_norm = norm; S->clip_coef = coef; S->step = t; S->bc1 = bc1; S->bc2 = bc2;
    mypad[kPadCounter] = 0u;
  }
  if (threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + kPadDone + d.rank, epoch);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. wait for every peer's shadow / parameter writes, then advance the local epoch
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) dp_wait_kernel(DpParams d) {
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  dp_wait_all(mypad, kPadDone, d.world, epoch, &d.a.state->err);
  if (threadIdx.x == 0) mypad[kPadEpoch] = epoch;
}

}  // namespace fnd
