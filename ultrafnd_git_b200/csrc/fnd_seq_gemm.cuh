// fnd_seq_gemm.cuh — persistent, throughput-oriented tcgen05 GEMM for the sequence front-end (Tier B).
//
//   C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ R[M,N]) (-> GELU)        bf16 operands, fp32 accumulation in TMEM
//
// Where the grouped kernel of fnd_gemm.cuh is built for LATENCY (batch-128 chains: narrow tiles, one accumulator, one
// tile per CTA), the token-level projections of the sequence front-end are THROUGHPUT problems (M = batch x tokens up to
// 10^5 rows, N = K = d_model): here one CTA per SM stays resident and walks the tile list,
//   * tiles are 128 x BN with BN up to 256 (UMMA 128 x 256 x 16: one instruction keeps the tensor pipe busy for 128
//     cycles while reading 12 KB of shared memory — within the 128 B/clk shared-memory port),
//   * TWO accumulators live in TMEM (2 x BN columns, up to all 512), so the MMA warp starts tile i+1 while the four
//     epilogue warps drain tile i (tcgen05.ld -> bias / residual / activation -> bf16 into a swizzled staging buffer ->
//     TMA store; the residual slab is TMA-loaded into the same buffer),
//   * operands arrive by TMA (SWIZZLE_128B) through a 4..8-stage ring that runs ahead across tile boundaries,
//   * consecutive tile ids share the A panel (m-major order), so the co-resident CTAs re-read A from L2, not HBM.
// TMA zero-fills out-of-bounds rows, stores are row-masked: any M, any K that is a multiple of 8, N a multiple of 8.
//
// No counterpart in the reference; the nearest call sites are nn.Linear applications,
// e.g. src/models/fusion/cross_modal_transformer.py:147-150.
#pragma once
#include "fnd_common.cuh"

namespace fnd {

constexpr int kSeqGemmBM = 128;
constexpr int kSeqGemmBK = 64;
constexpr int kSeqGemmMaxStages = 8;
constexpr int kSeqGemmThreads = 192;             // warp0 TMA, warp1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int kSeqGemmHeader = 1024;
constexpr int kSeqGemmRingBudget = 192 * 1024;
constexpr int kSeqGemmStageOutBytes = kSeqGemmBM * 64 * 2;    // one 128 x 64 bf16 output slab
constexpr int kSeqGemmSmemMax = kSeqGemmRingBudget + 2 * kSeqGemmStageOutBytes + kSeqGemmHeader + 1024;   // residual buffers come out of the ring budget

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct alignas(64) SeqGemmParams {
  CUtensorMap tmA, tmB;
  CUtensorMap tmC, tmR;                          // bf16 output / residual, box 64 x 128 (valid when out_bf / resid are set)
  int M, N, K;
  int bn;                                        // 64 / 128 / 256
  int tiles_m, tiles_n, kblocks, nstages, stage_bytes;
  const float* bias;                             // [N] or null
  const __nv_bfloat16* resid;                    // [M, resid_pitch] or null
  int resid_pitch;
  int act;                                       // 1 = exact-erf GELU
  __nv_bfloat16* out_bf;                         // [M, out_pitch] or null
  int out_pitch;
  float* out_f32;                                // [M, f32_pitch] or null
  int f32_pitch;
  int* err;
  // Weight-gradient mode (mn = 1): C[M,N] = sum_k A[k,m] B[k,n] with BOTH operands stored [K][rows] (rows contiguous) — dW =
  // dY^T X read token-major in place (MN-major UMMA descriptors, 64 x 64 TMA boxes), fp32 output. The token reduction is cut
  // into `splits` ranges of kb_per_split k-blocks; split s writes its partial tile to out_f32 + s * split_stride (summed in
  // split order by seq_reduce_partials_kernel: deterministic), which turns 32..192 output tiles into >= 4 waves of items.
  int mn;
  int splits, kb_per_split;
  long long split_stride;                        // floats between the partial outputs of consecutive splits
  long long* dbg;                                // probe aid (fnd_seq_debug_gemm_stamps, pair kernel): 8 cycle counters per cluster
  int dbg_mmas;                                  // probe aid (FND_SEQ_DBG_MMAS, pair kernel): issue only this many of the 4 UMMAs per k-block
};

__global__ void __launch_bounds__(kSeqGemmThreads, 1) seq_gemm_kernel(const __grid_constant__ SeqGemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kSeqGemmMaxStages;
  uint64_t* tfull_bar = empty_bar + kSeqGemmMaxStages;     // [2] accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2] accumulator drained
  uint64_t* rfull_bar = tempty_bar + 2;                    // [2] residual slab landed in the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + 2);
  uint8_t* ring = smem + kSeqGemmHeader;
  uint8_t* stage = ring + P.nstages * P.stage_bytes;       // 2 x 16 KB output staging (128 rows x 64 bf16, SWIZZLE_128B)
  uint8_t* rbuf = stage + 2 * kSeqGemmStageOutBytes;       // 2 x 16 KB residual slabs (only when P.resid; host sizes the ring)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, nstages = P.nstages, stage_bytes = P.stage_bytes, kblocks = P.kblocks;
  const int ntiles = P.tiles_m * P.tiles_n;
  const int nitems = ntiles * P.splits;            // work item = (k-range split, tile); splits == 1 outside the wgrad mode
  const bool mn = P.mn != 0;
  const uint32_t tmem_cols = static_cast<uint32_t>(2 * bn);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    if (P.out_bf) tma_prefetch_desc(&P.tmC);
    if (P.resid) tma_prefetch_desc(&P.tmR);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); mbar_init(&rfull_bar[a], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: runs ahead across tile boundaries =================
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(stage_bytes);
      int s = 0;
      uint32_t ph = 1u;                          // parity to wait for on empty_bar (fresh barrier: passes)
      bool ok = true;
#pragma unroll 1
      for (int item = blockIdx.x; item < nitems && ok; item += gridDim.x) {
        const int split = item / ntiles, tile = item - split * ntiles;
        const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
        const int am = tm * kSeqGemmBM, bnr = tn * bn;
        const int kb0 = split * P.kb_per_split, kb1 = min(kblocks, kb0 + P.kb_per_split);
#pragma unroll 1
        for (int kb = kb0; kb < kb1; ++kb) {
          ok = mbar_wait_fast(&empty_bar[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
          if (!ok) break;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          uint8_t* sA = ring + s * stage_bytes;
          uint8_t* sB = sA + kSeqGemmBM * kSeqGemmBK * 2;
          if (!mn) {
            tma_load_2d(sA, &P.tmA, &full_bar[s], kb * kSeqGemmBK, am, kEvictNormal);
            tma_load_2d(sB, &P.tmB, &full_bar[s], kb * kSeqGemmBK, bnr, kEvictLast);
          } else {
            // 64 (rows, contiguous) x 64 (k) boxes: one 128-byte swizzle atom wide, 8 KB each, LBO = 8192 between atoms
            tma_load_2d(sA, &P.tmA, &full_bar[s], am, kb * kSeqGemmBK, kEvictNormal);
            tma_load_2d(sA + 8192, &P.tmA, &full_bar[s], am + 64, kb * kSeqGemmBK, kEvictNormal);
            for (int ch = 0; ch < bn / 64; ++ch)
              tma_load_2d(sB + ch * 8192, &P.tmB, &full_bar[s], bnr + ch * 64, kb * kSeqGemmBK, kEvictNormal);
          }
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (lean loop: incremental stage / parity, 32-bit descriptor arithmetic; warp-uniform
    // control flow with ONE elected lane around the issue, so operands reach the uniform datapath without
    // divergence-safe conversion loops) =================
    {
      const uint32_t idesc = make_idesc_bf16(kSeqGemmBM, bn, mn ? 1 : 0, mn ? 1 : 0);
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(ring), mn ? 8192 : 16);
      const uint32_t b_off = (kSeqGemmBM * kSeqGemmBK * 2) >> 4;
      const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
      // K-major: 16 contraction elements = 32 B inside the swizzled row; MN-major: 16 contraction rows of 128 B = 2048 B
      const uint32_t kstep = mn ? 128u : 2u;
      int s = 0, lt = 0;
      uint32_t ph = 0u;
      bool ok = true;
#pragma unroll 1
      for (int item = blockIdx.x; item < nitems && ok; item += gridDim.x, ++lt) {
        const int split = item / ntiles;
        const int kb0 = split * P.kb_per_split, kb1 = min(kblocks, kb0 + P.kb_per_split);
        const int ab = lt & 1;
        const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
        ok = mbar_wait_fast(&tempty_bar[ab], aph ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(ab * bn);
#pragma unroll 1
        for (int kb = kb0; kb < kb1; ++kb) {
          ok = mbar_wait_fast(&full_bar[s], ph, P.err, FND_DEV_TIMEOUT_MMA);
          if (!ok) break;
          tc_fence_after_sync();
          const uint32_t al = a_lo0 + static_cast<uint32_t>(s) * stage_step;
          const uint32_t bl = al + b_off;
          if (elect_one()) {
            umma_f16(tacc, desc64(al, dhi), desc64(bl, dhi), idesc, kb != kb0 ? 1u : 0u);
            umma_f16(tacc, desc64(al + kstep, dhi), desc64(bl + kstep, dhi), idesc, 1u);
            umma_f16(tacc, desc64(al + 2 * kstep, dhi), desc64(bl + 2 * kstep, dhi), idesc, 1u);
            umma_f16(tacc, desc64(al + 3 * kstep, dhi), desc64(bl + 3 * kstep, dhi), idesc, 1u);
            umma_commit(&empty_bar[s]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[ab]);
          }
          __syncwarp();
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quarter = warp % 4 =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    int lt = 0;
    uint32_t slab_ctr = 0;                                   // 64-column slabs staged so far (selects the staging buffer)
    // Residual slabs are prefetched TWO slabs ahead into their own buffers (they do not depend on the accumulator, so
    // the first two are requested before the first tile's MMAs have even finished): a load issued on demand cost a full
    // L2/HBM round trip per slab and made the epilogue, not the MMA, the pace of the residual GEMMs.
    int pf_tile = blockIdx.x, pf_c = 0;
    uint32_t pf_n = 0;
    auto pf_issue = [&]() {                                 // (residuals never occur in the wgrad mode: items == tiles here)
      if (pf_tile >= ntiles) return;
      const int ptm = pf_tile / P.tiles_n, ptn = pf_tile - ptm * P.tiles_n;
      const uint32_t b = pf_n & 1u;
      mbar_arrive_expect_tx(&rfull_bar[b], kSeqGemmStageOutBytes);
      tma_load_2d(rbuf + b * kSeqGemmStageOutBytes, &P.tmR, &rfull_bar[b], ptn * bn + pf_c, ptm * kSeqGemmBM, kEvictFirst);
      ++pf_n;
      pf_c += 64;
      if (pf_c >= bn || ptn * bn + pf_c >= P.N) { pf_c = 0; pf_tile += gridDim.x; }
    };
    if (P.resid && P.out_bf && epi_tid == 0) { pf_issue(); pf_issue(); }
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++lt) {
      const int split = item / ntiles, tile = item - split * ntiles;
      const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
      const int ab = lt & 1;
      const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
      const bool ok = mbar_wait(&tfull_bar[ab], aph, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      const int m0 = tm * kSeqGemmBM;
      const int m = m0 + row;
      const uint32_t taddr = tmem_base + lane_addr + static_cast<uint32_t>(ab * bn);
      const int nb = tn * bn;
      if (P.out_bf) {
        // ---- bf16 output: 64-column slabs go through a swizzled shared-memory staging buffer and leave by TMA store.
        // A thread owns an accumulator ROW; storing rows straight from registers would make every warp-level store touch
        // 32 different 128-byte lines (LSU-bound: measured 20 K cycles per 128x256 tile against 8 K cycles of MMA). The
        // residual slab arrives the same way (TMA load into the staging buffer, read back by the owning thread). ----
#pragma unroll 1
        for (int c = 0; c < bn; c += 64) {
          const int n0 = nb + c;
          if (n0 >= P.N) break;                              // tile-uniform
          const uint32_t sb = slab_ctr & 1u;
          uint8_t* stg = stage + sb * kSeqGemmStageOutBytes;
          if (epi_tid == 0) tma_store_wait_read<1>();        // the store issued two slabs ago has drained this buffer
          epi_bar_sync();
          if (P.resid) mbar_wait(&rfull_bar[sb], (slab_ctr >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
          const uint8_t* rrow = rbuf + sb * kSeqGemmStageOutBytes + row * 128;
          // the slab's 64 bias values, requested back to back BEFORE the accumulator load is waited for: loads left next to
          // their use were issued one chunk at a time (ncu: the epilogue's top stalls were the FADDs behind them)
          float4 bq[16];
          if (P.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(P.bias + min(n0 + 4 * i, P.N - 4)));
          }
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c, r0);
          tmem_ld_32x32(taddr + c + 32, r1);
          tmem_ld_wait();
          uint8_t* prow = stg + row * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {                   // 8 chunks of 8 columns
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(ch < 4 ? r0[ch * 8 + j] : r1[(ch - 4) * 8 + j]);
            if (P.bias) {                                    // (columns >= N are clipped by the TMA store)
              const float4 b0 = bq[2 * ch], b1 = bq[2 * ch + 1];
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            uint4* slot = reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4));
            if (P.resid) {
              const uint4 u = *reinterpret_cast<const uint4*>(rrow + ((ch ^ (row & 7)) << 4));
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
                v[2 * t] += __low2float(h2);
                v[2 * t + 1] += __high2float(h2);
              }
            }
            if (P.act == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
            }
            *slot = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
          fence_proxy_async_smem();
          epi_bar_sync();
          if (epi_tid == 0) {
            if (ok) {
              tma_store_2d(&P.tmC, stg, n0, m0);             // rows >= M and columns >= N are clipped by the tensor map
              tma_store_commit();
            }
            if (P.resid) pf_issue();                         // every thread has read this residual buffer: refill it
          }
          ++slab_ctr;
        }
      }
      if (P.out_f32) {
        // ---- fp32 output (the small pooled-head GEMMs): direct row stores ----
        const bool row_ok = ok && m < P.M;
#pragma unroll 1
        for (int c = 0; c < bn; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          const int n0 = nb + c;
          if (!row_ok || n0 >= P.N) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const int ncols = min(32, P.N - n0);              // multiple of 8
          if (P.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
          }
          if (P.resid) {
            const __nv_bfloat16* rp = P.resid + static_cast<size_t>(m) * P.resid_pitch + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < ncols) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + j));
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
                  v[j + 2 * t] += __low2float(h2);
                  v[j + 2 * t + 1] += __high2float(h2);
                }
              }
            }
          }
          if (P.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          float* op = P.out_f32 + static_cast<size_t>(split) * P.split_stride + static_cast<size_t>(m) * P.f32_pitch + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      // every column of this accumulator is in registers / staged: hand it back to the MMA warp
      tc_fence_before_sync();
      mbar_arrive(&tempty_bar[ab]);
    }
    if (epi_tid == 0) tma_store_wait<0>();                   // all output tiles written before the CTA retires
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace fnd
