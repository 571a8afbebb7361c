// fnd_dp.cuh — data-parallel optimizer step over NVLink peer memory (one process per GPU, batch-sharded replicas).
//
// The reference trains on one device (src/training/forensic_trainer.py:285-298); its clip_grad_norm_ + AdamW pair is
// what this file replaces when the batch is sharded over N GPUs. Instead of "NCCL all-reduce of the 51 MB fp32 gradient,
// then N identical AdamW passes over 357 MB each", the gradient exchange, the clip and the optimizer are ONE sharded
// sequence over symmetric (peer-mapped) buffers. Every rank owns 1/N of each of three arena ranges (its "slice"):
//
//   dp_push_kernel     rank r stores, for every peer p, the part of ITS gradient that p owns into p's staging buffer
//                      (slot r) with 128-bit P2P stores over NVLink/NVSwitch — the reduce-scatter's data movement,
//                      sender-driven so that no handshake precedes it. It runs twice per step: for the fuse_mlp.0 /
//                      fuse_mlp.3 weight gradients (70 % of all gradient bytes), which are complete two thirds of the way
//                      through the backward pass and are pushed from a side stream UNDER the rest of the backward (one
//                      small CTA per SM, so the backward's GEMM CTAs still fit), and for everything else at the end.
//                      In bf16 mode the pieces travel as bf16 (the sum is taken in fp32);
//   dp_reduce_kernel   rank r sums its own slice and the N-1 staged pieces in rank order (local memory only), keeps the
//                      reduced slice and publishes the slice's sum of squares to every peer;
//   dp_adamw_kernel    every rank adds the N partial sums in rank order (bit-identical clip coefficient everywhere),
//                      runs AdamW on its slice only (1/N of the 357 MB optimizer stream; fp32 master, m and v stay
//                      sharded, ZeRO-1 style) and writes the refreshed bf16 operand shadows — the only copy of the
//                      weights the forward/backward kernels read — plus the small fp32 parameters (biases, gates, ...)
//                      into EVERY rank's buffers with P2P stores (the all-gather);
//   dp_wait_kernel     blocks the stream until every peer's shadow writes have landed here.
//
// Deferred update (optional): 70 % of the all-gather bytes belong to fuse_mlp.0 / fuse_mlp.3, whose weights are first
// read by the FIFTH kernel of the next forward pass. With deferral the optimizer step above covers only the other two
// ranges; the fuse_mlp slice is updated and all-gathered by a second dp_adamw_kernel launch that the NEXT training step
// starts on a side stream and joins right before gemm_fuse0 — the all-gather runs under prep / projections / q-k-v /
// assemble instead of on the critical path. fnd_dp_flush applies a pending deferred update immediately (before an
// evaluation pass, a learning-rate change or a state_dict read).
//
// Measured on this pool's 8 x B200 NVSwitch box (tools/p2p_probe.py, all ranks active): P2P stores 548 GB/s per rank
// and direction, P2P loads 480 GB/s, multimem.ld_reduce / multimem.st 550 GB/s on the limiting direction, copy-engine
// pushes 198 GB/s — hence SM stores for the data movement, and pushes (not pulls) so the early part needs no handshake.
//
// Cross-GPU ordering uses epoch flags in each rank's symmetric comm pad (st.release.sys / ld.acquire.sys): a flag
// holds the number of the last step for which the event happened, so nothing is ever reset and a captured CUDA graph
// can be replayed. The kernels only ever wait for events that peers produce without needing anything further from
// this rank, so the sequence cannot deadlock as long as every rank runs the same steps.
#pragma once
#include "fnd_optim.cuh"

namespace fnd {

constexpr int kDpMaxWorld = 8;
constexpr int kDpMaxSeg = 3;
// comm pad layout (uint32 words): [0,8) early pieces pushed (epochs) | [8,16) late pieces pushed | [16,24) partial
// norms ready | [24,32) shadows written | [32,40) float partial sums of squares | [40] local epoch counter |
// [41] CTA counter of the main-stream kernels | [42] CTA counter of the early (side-stream) push
// [43] pending: epoch whose deferred (range 0) update has not been applied yet, 0 = none | [48,56) deferred shadows written
constexpr int kPadReadyEarly = 0, kPadReadyLate = 8, kPadPartialReady = 16, kPadDone = 24, kPadPartial = 32, kPadEpoch = 40,
              kPadCounter = 41, kPadCounterEarly = 42, kPadPending = 43, kPadDoneDeferred = 48;
constexpr int kPadWords = 64;

struct DpParams {
  int rank, world;
  const float* grads;                     // this rank's gradient arena
  float* params[kDpMaxWorld];             // fp32 parameter arena of every rank (peer-mapped)
  __nv_bfloat16* sh_hi[kDpMaxWorld];      // bf16 operand shadows of every rank
  __nv_bfloat16* sh_lo[kDpMaxWorld];      // residual planes (fp32x3 mode) or null
  unsigned int* pad[kDpMaxWorld];         // comm pad of every rank
  // NVSwitch multicast mappings of the same buffers (one store lands on every rank; null when the fabric has no multicast)
  float* mc_params;
  __nv_bfloat16* mc_sh_hi;
  __nv_bfloat16* mc_sh_lo;
  const float* mc_grads;                  // multicast view of the gradient arenas (pull mode: in-switch reduction)
  const __nv_bfloat16* mc_grads_bf;       // ... and of their bf16 mirrors (null: reduce the fp32 gradients)
  void* stage[kDpMaxWorld];               // staging buffer of every rank: [world slots][slot_cap] fp32 or bf16
  size_t slot_cap;                        // elements per staging slot (>= the largest slice)
  int stage_bf16;                         // 1: pieces travel as bf16
  float* gred;                            // local: reduced gradient slice, segments back to back
  float* slots;                           // local: per-CTA sums of squares of dp_reduce_kernel
  // Slices: rank p owns [seg_lo[p][s], seg_hi[p][s]) of arena range s. Range 0 is the "early" one (fuse_mlp.0.weight and
  // fuse_mlp.3.weight), ranges 1 and 2 the arena before and after it. Inside a staging slot / gred the three pieces of a
  // slice lie back to back at seg_goff[p][s].
  int nseg;
  size_t seg_lo[kDpMaxWorld][kDpMaxSeg], seg_hi[kDpMaxWorld][kDpMaxSeg], seg_goff[kDpMaxWorld][kDpMaxSeg];
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
// 1a. push: this rank's contributions to every peer's slice -> that peer's staging slot for this rank
// ---------------------------------------------------------------------------------------------------------------
// Segments [s0, s1); `bank` is the flag bank raised on every peer once all stores of this launch are globally visible;
// `ctr` the pad word used to elect the last CTA (the early launch runs concurrently with main-stream kernels).
template <bool BF16>
__global__ void __launch_bounds__(256, 4) dp_push_kernel(DpParams d, int s0, int s1, int bank, int ctr) {
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int sg = s0; sg < s1; ++sg) {
    for (int q = 1; q < d.world; ++q) {
      const int p = (d.rank + q) % d.world;            // start with a different peer on every rank: spreads the links
      const size_t lo = d.seg_lo[p][sg], n4 = (d.seg_hi[p][sg] - lo) >> 2;
      const float* src = d.grads + lo;
      const size_t doff = static_cast<size_t>(d.rank) * d.slot_cap + d.seg_goff[p][sg];
      // four independent 128-bit loads in flight per thread; the stores are fire-and-forget
      for (size_t i4 = tid; i4 < n4; i4 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i4 + u * stride < n4) v[u] = __ldcg(reinterpret_cast<const float4*>(src) + i4 + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t j4 = i4 + u * stride;
          if (j4 < n4) {
            if (BF16) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.stage[p]) + doff;
              *reinterpret_cast<uint2*>(dst + j4 * 4) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
            } else {
              float* dst = static_cast<float*>(d.stage[p]) + doff;
              *reinterpret_cast<float4*>(dst + j4 * 4) = v[u];
            }
          }
        }
      }
    }
  }
  // every P2P store of this CTA is ordered before its counter bump; the last CTA raises the flag on every peer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(mypad + ctr, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) mypad[ctr] = 0u;
  if (threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + bank + d.rank, epoch);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 1a'. residual push of the FUSED mode: the weight-gradient GEMM's epilogue has already stored every GEMM-weight tile
//      into its owner's staging slot (fnd_gemm.cuh, DpRoute); what is left are the two arena intervals that launch does
//      not route — pre.0.weight (row pitch 514: not 16-byte aligned) and the small non-GEMM parameters behind the
//      shadows (biases, gates, thresholds, leaf tables, evidence MLPs; written by the finalize CTAs) — about 2 % of the
//      gradient. They are pushed here, to EVERY owner including this rank itself (the fused reduce reads all N pieces
//      from staging), as bf16; the last CTA then raises the "late" flags, which also cover the epilogue's stores (that
//      launch has completed: ordinary stream order).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) dp_push_residual_kernel(DpParams d, size_t lo0, size_t hi0, size_t lo1, size_t hi1) {
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int sg = 0; sg < d.nseg; ++sg) {
    for (int q = 0; q < d.world; ++q) {
      const int p = (d.rank + q) % d.world;
      for (int iv = 0; iv < 2; ++iv) {
        const size_t a = iv ? lo1 : lo0, b = iv ? hi1 : hi0;
        const size_t lo = d.seg_lo[p][sg] > a ? d.seg_lo[p][sg] : a;
        const size_t hi = d.seg_hi[p][sg] < b ? d.seg_hi[p][sg] : b;
        if (hi <= lo) continue;
        const size_t n4 = (hi - lo) >> 2;                        // interval and segment bounds are multiples of 64
        const float* src = d.grads + lo;
        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(d.stage[p]) + static_cast<size_t>(d.rank) * d.slot_cap + d.seg_goff[p][sg] +
                             (lo - d.seg_lo[p][sg]);
        for (size_t i4 = tid; i4 < n4; i4 += stride) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i4);
          *reinterpret_cast<uint2*>(dst + i4 * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(mypad + kPadCounter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) mypad[kPadCounter] = 0u;
  if (threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + kPadReadyLate + d.rank, epoch);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 1'. pull: the whole reduce-scatter as ONE kernel when the fabric has multicast — every rank reads its slice through
//     the NVSwitch with multimem.ld_reduce, which fetches the same address from all N gradient arenas and returns the
//     sum (accumulated in fp32 inside the switch). No staging, no sender-side kernel, one flag round. GEMM-weight
//     gradients are read from the bf16 mirror when there is one (half the wire bytes; the sum comes back rounded to
//     bf16), everything else — and everything in fp32 mode — from the fp32 arenas.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dp_pull_kernel(DpParams d) {
  __shared__ float red[8];
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  // all of this rank's gradients are complete (stream order): tell every peer, then wait for all of them
  if (blockIdx.x == 0 && threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + kPadReadyLate + d.rank, epoch);
  }
  dp_wait_all(mypad, kPadReadyLate, d.world, epoch, &d.a.state->err);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t rp0 = d.a.rp_begin, rp1 = d.a.rp_end, nsh = d.a.n_shadow;
  float ss = 0.f;
  for (int sg = 0; sg < d.nseg; ++sg) {
    const size_t lo = d.seg_lo[d.rank][sg], n8 = (d.seg_hi[d.rank][sg] - lo) >> 3;
    float* out = d.gred + d.seg_goff[d.rank][sg];
    for (size_t i8 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i8 < n8; i8 += stride) {
      const size_t i = lo + i8 * 8;
      float v[8];
      // mirrored: a GEMM weight other than pre.0.weight (tensor boundaries are multiples of 64 elements)
      const bool mirrored = d.mc_grads_bf != nullptr && i < nsh && !(i >= rp0 && i < rp1);
      if (mirrored) {
        uint32_t r0, r1, r2, r3;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(d.mc_grads_bf + i) : "memory");
        const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[q]);
          v[2 * q] = __low2float(h2); v[2 * q + 1] = __high2float(h2);
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v[4 * h]), "=f"(v[4 * h + 1]), "=f"(v[4 * h + 2]), "=f"(v[4 * h + 3])
                       : "l"(d.mc_grads + i + 4 * h) : "memory");
      }
      *reinterpret_cast<float4*>(out + i8 * 8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + i8 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
      for (int q = 0; q < 8; ++q) ss = fmaf(v[q], v[q], ss);
    }
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
// 1b. reduce: own slice + the staged pieces of every peer, in rank order (local memory only) + slice norm
// ---------------------------------------------------------------------------------------------------------------
// `early_used`: segment 0 was pushed by a separate (early) launch with its own flag bank.
// `own_from_stage` (fused mode): this rank's own piece lies in its staging slot too (bf16), not in the gradient arena.
template <bool BF16>
__global__ void __launch_bounds__(256) dp_reduce_kernel(DpParams d, int early_used, int own_from_stage) {
  __shared__ float red[8];
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  if (early_used) dp_wait_all(mypad, kPadReadyEarly, d.world, epoch, &d.a.state->err);   // (own flags: raised by the own pushes)
  dp_wait_all(mypad, kPadReadyLate, d.world, epoch, &d.a.state->err);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  float ss = 0.f;
  for (int sg = 0; sg < d.nseg; ++sg) {
    const size_t lo = d.seg_lo[d.rank][sg], n4 = (d.seg_hi[d.rank][sg] - lo) >> 2;
    const size_t goff = d.seg_goff[d.rank][sg];
    float* out = d.gred + goff;
    for (size_t i4 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
      float4 t[kDpMaxWorld];
#pragma unroll
      for (int p = 0; p < kDpMaxWorld; ++p) {
        if (p < d.world) {
          if (p == d.rank && !own_from_stage) {
            t[p] = __ldcg(reinterpret_cast<const float4*>(d.grads + lo) + i4);
          } else if (BF16) {
            const __nv_bfloat16* sp = static_cast<const __nv_bfloat16*>(d.stage[d.rank]) + static_cast<size_t>(p) * d.slot_cap + goff;
            const uint2 raw = __ldcg(reinterpret_cast<const uint2*>(sp) + i4);
            const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), b2 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
            t[p] = make_float4(__low2float(a2), __high2float(a2), __low2float(b2), __high2float(b2));
          } else {
            const float* sp = static_cast<const float*>(d.stage[d.rank]) + static_cast<size_t>(p) * d.slot_cap + goff;
            t[p] = __ldcg(reinterpret_cast<const float4*>(sp) + i4);
          }
        }
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kDpMaxWorld; ++p)      // summed in rank order
        if (p < d.world) { acc.x += t[p].x; acc.y += t[p].y; acc.z += t[p].z; acc.w += t[p].w; }
      *reinterpret_cast<float4*>(out + i4 * 4) = acc;
      ss += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
    }
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
__device__ __forceinline__ void multimem_st16(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

// Eight consecutive updated parameters [i, i+8) -> every rank: bf16 shadow (one 128-bit store per destination), or the
// fp32 values themselves for the small parameters the kernels read from the arena.
__device__ __forceinline__ void dp_publish8(const DpParams& d, size_t i, const float (&x)[8]) {
  const AdamWParams& a = d.a;
  if (i < a.n_shadow) {
    uint4 h, l = make_uint4(0, 0, 0, 0);
    h.x = pack_bf16x2(x[0], x[1]); h.y = pack_bf16x2(x[2], x[3]); h.z = pack_bf16x2(x[4], x[5]); h.w = pack_bf16x2(x[6], x[7]);
    if (a.sh_lo) {
      const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
      uint32_t ll[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&hh[q]);
        ll[q] = pack_bf16x2(x[2 * q] - __low2float(h2), x[2 * q + 1] - __high2float(h2));
      }
      l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
    if (d.mc_sh_hi) {
      // one multicast store: the switch replicates it into every rank's shadow (this rank's included)
      multimem_st16(d.mc_sh_hi + i, h);
      if (a.sh_lo) multimem_st16(d.mc_sh_lo + i, l);
    } else {
#pragma unroll
      for (int q = 0; q < kDpMaxWorld; ++q) {
        if (q < d.world) {
          const int p = (d.rank + q) % d.world;           // own copy first, then a different peer order on every rank
          *reinterpret_cast<uint4*>(d.sh_hi[p] + i) = h;
          if (a.sh_lo) *reinterpret_cast<uint4*>(d.sh_lo[p] + i) = l;
        }
      }
    }
  } else if (d.mc_params) {
    // (rewrites this rank's own copy with the same values)
    multimem_st16(d.mc_params + i, make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])));
    multimem_st16(d.mc_params + i + 4, make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
  } else {
#pragma unroll
    for (int q = 1; q < kDpMaxWorld; ++q) {
      if (q < d.world) {
        const int p = (d.rank + q) % d.world;
        *reinterpret_cast<float4*>(d.params[p] + i) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(d.params[p] + i + 4) = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
  }
  if (i + 8 > a.rp_begin && i < a.rp_end) {
    // re-pitched pre.0.weight shadow (+ its fp32 aux columns, which the epilogue reads from the arena)
    const size_t rp_off = static_cast<size_t>(a.rp_hi - a.sh_hi);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const size_t e = i + q;
      if (e >= a.rp_begin && e < a.rp_end) {
        const size_t r = (e - a.rp_begin) / a.rp_cols, c = (e - a.rp_begin) % a.rp_cols;
        __nv_bfloat16 hb, lb;
        split_bf16(x[q], hb, lb);
        for (int p = 0; p < d.world; ++p) {
          d.sh_hi[p][rp_off + r * a.rp_pitch + c] = hb;
          if (a.sh_lo) d.sh_lo[p][rp_off + r * a.rp_pitch + c] = lb;
          if (c >= static_cast<size_t>(a.rp_cols - 2) && p != d.rank) d.params[p][e] = x[q];
        }
      }
    }
  }
}

// mode 0: the step's optimizer launch over segments [s0, s1): waits for the partial norms, forms the clip coefficient,
//         publishes norm / coefficient / step, raises kPadDone; when s0 > 0 it marks range 0 as pending (deferred).
// mode 1: the deferred launch over segment 0 (next step's side stream, or fnd_dp_flush): no-op unless pending; uses the
//         coefficient and bias corrections published by the mode-0 launch of the same optimizer step; raises
//         kPadDoneDeferred (always) and clears the pending mark. `ctr` = pad word used to elect the last CTA.
__global__ void __launch_bounds__(256) dp_adamw_kernel(DpParams d, int s0, int s1, int mode, int ctr) {
  __shared__ float s_norm;
  __shared__ int is_last;
  unsigned int* mypad = d.pad[d.rank];
  DevState* S = d.a.state;
  unsigned int epoch;
  float norm = 0.f, coef, bc1, bc2;
  int t = 0;
  const float lr = S->lr, b1 = S->beta1, b2 = S->beta2, eps = S->eps;
  if (mode == 0) {
    epoch = mypad[kPadEpoch] + 1u;
    dp_wait_all(mypad, kPadPartialReady, d.world, epoch, &S->err);
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int p = 0; p < d.world; ++p) tot += static_cast<double>(reinterpret_cast<volatile float*>(mypad)[kPadPartial + p]);
      s_norm = static_cast<float>(sqrt(tot));
    }
    __syncthreads();
    norm = s_norm;
    coef = clip_coef_of(S->max_norm, norm);
    t = S->step + 1;
    bc1 = 1.0f - powf(b1, static_cast<float>(t));
    bc2 = 1.0f - powf(b2, static_cast<float>(t));
  } else {
    // kPadDoneDeferred flags mean "this rank has no deferred update pending for epochs <= value": raise them with the
    // last completed epoch whether or not there was work (every rank runs this launch at the same points)
    epoch = mypad[kPadEpoch];
    if (mypad[kPadPending] == 0u) {          // nothing deferred (first step, or already flushed): uniform over the grid
      if (blockIdx.x == 0 && threadIdx.x < static_cast<unsigned>(d.world))
        st_release_sys(d.pad[threadIdx.x] + kPadDoneDeferred + d.rank, epoch);
      return;
    }
    coef = S->clip_coef; bc1 = S->bc1; bc2 = S->bc2;
  }
  const float decay = 1.0f - lr * S->weight_decay;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const AdamWParams& a = d.a;
  const uint64_t pol = l2_policy_evict_first();      // optimizer streams must not push the weights / activations out of L2
  auto update4 = [&](size_t i, const float* gsrc4, float (&out)[4]) {
    float4 p = ld_f4_policy(a.p + i, pol);
    const float4 g4 = __ldcg(reinterpret_cast<const float4*>(gsrc4));
    float4 m = ld_f4_policy(a.m + i, pol);
    float4 v = ld_f4_policy(a.v + i, pol);
    float* pp = &p.x; float* mp = &m.x; float* vp = &v.x; const float* gp = &g4.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gq = gp[q] * coef;
      pp[q] *= decay;
      mp[q] = b1 * mp[q] + (1.0f - b1) * gq;
      vp[q] = b2 * vp[q] + (1.0f - b2) * gq * gq;
      const float denom = sqrtf(vp[q]) * inv_sqrt_bc2 + eps;
      pp[q] -= step_size * (mp[q] / denom);
      out[q] = pp[q];
    }
    st_f4_policy(a.p + i, p, pol);
    st_f4_policy(a.m + i, m, pol);
    st_f4_policy(a.v + i, v, pol);
  };
  for (int sg = s0; sg < s1; ++sg) {
    const size_t len = d.seg_hi[d.rank][sg] - d.seg_lo[d.rank][sg];
    const size_t n8 = len >> 3;                                    // slices are multiples of 1024 elements ...
    const bool tail4 = (len & 7) != 0;                             // ... except the last rank's (a multiple of 4)
    const float* gsrc = d.gred + d.seg_goff[d.rank][sg];
    for (size_t i8 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i8 < n8;
         i8 += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t i = d.seg_lo[d.rank][sg] + i8 * 8;
      float lo4[4], hi4[4];
      update4(i, gsrc + i8 * 8, lo4);
      update4(i + 4, gsrc + i8 * 8 + 4, hi4);
      const float x[8] = {lo4[0], lo4[1], lo4[2], lo4[3], hi4[0], hi4[1], hi4[2], hi4[3]};
      dp_publish8(d, i, x);
    }
    if (tail4 && blockIdx.x == 0 && threadIdx.x == 0) {
      const size_t i = d.seg_lo[d.rank][sg] + n8 * 8;
      float t4[4];
      update4(i, gsrc + n8 * 8, t4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {                                // element-wise publication of the 4-element tail
        const size_t e = i + q;
        for (int pr = 0; pr < d.world; ++pr) {
          if (e < a.n_shadow) {
            __nv_bfloat16 hb, lb;
            split_bf16(t4[q], hb, lb);
            d.sh_hi[pr][e] = hb;
            if (a.sh_lo) d.sh_lo[pr][e] = lb;
          } else if (pr != d.rank) {
            d.params[pr][e] = t4[q];
          }
        }
      }
    }
  }
  // every P2P store of this CTA is ordered before its counter bump; the last CTA tells the peers
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(mypad + ctr, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) {
    if (mode == 0) {
      S->grad_norm = norm; S->clip_coef = coef; S->step = t; S->bc1 = bc1; S->bc2 = bc2;
      mypad[kPadPending] = s0 > 0 ? epoch : 0u;
    } else {
      mypad[kPadPending] = 0u;
    }
    mypad[ctr] = 0u;
  }
  if (threadIdx.x < static_cast<unsigned>(d.world)) {
    __threadfence_system();
    st_release_sys(d.pad[threadIdx.x] + (mode == 0 ? kPadDone : kPadDoneDeferred) + d.rank, epoch);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. wait for every peer's shadow / parameter writes, then advance the local epoch
// ---------------------------------------------------------------------------------------------------------------
// deferred = 0: end of an optimizer step (waits kPadDone for the running epoch, then advances the local epoch);
// deferred = 1: after a deferred launch (waits until every peer's deferred shadows of the LAST COMPLETED epoch are here;
//               trivially satisfied when nothing was deferred, because every rank defers or flushes in lockstep).
__global__ void __launch_bounds__(32) dp_wait_kernel(DpParams d, int deferred) {
  unsigned int* mypad = d.pad[d.rank];
  if (deferred) {
    dp_wait_all(mypad, kPadDoneDeferred, d.world, mypad[kPadEpoch], &d.a.state->err);
    return;
  }
  const unsigned int epoch = mypad[kPadEpoch] + 1u;
  dp_wait_all(mypad, kPadDone, d.world, epoch, &d.a.state->err);
  if (threadIdx.x == 0) mypad[kPadEpoch] = epoch;
}

}  // namespace fnd
