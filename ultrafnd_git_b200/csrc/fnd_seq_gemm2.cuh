// fnd_seq_gemm2.cuh — CTA-pair variant of the persistent sequence GEMM (fnd_seq_gemm.cuh): 256 x 256 tiles on TWO SMs.
//
//   C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ R[M,N]) (-> GELU)        bf16 operands, fp32 accumulation in TMEM, bf16 output
//
// What it buys, measured (tools/seq_probe.py, tools/gemm2_stamps.py, B200, stress shapes): each CTA stages its own 128 rows
// of A and only HALF of the W tile (32 KB per 64-deep k-block instead of 48 KB: a third less L2 -> SM traffic and UMMA
// operand-read bandwidth per flop) and receives its own 128 accumulator rows in its own TMEM, so the epilogue (tcgen05.ld
// -> bias / residual / GELU -> swizzled staging -> TMA store, residual slabs prefetched two ahead) is the single-CTA one.
// Kernel time is within 1-3 % of the single-CTA kernel (in_proj 32768 x 3072 x 1024: 160.7 vs 161.6 us; out_proj + residual
// 81.2 vs 83.5 us) and the co-attention block, where the GEMMs share the chip with the other stream's kernels, gains 1 %.
// Neither kernel is bound by operand traffic: cuBLAS runs the same shapes in 157.4 / 87.5 / 71.7 us against 161.0 / 93.4 /
// 69.6 us here — at K = 1024 the chip's 1000 W power cap and the epilogue's work per flop set the rate (an MMA-only loop of
// this kernel with neither loads nor epilogue reaches 1590 TFLOP/s, the same as the library's 8192^3 burst figure; adding
// the loads costs 17 us, the epilogue 15-30 us of the 160).
//
// Protocol (barriers sit at the same shared-memory offsets in both CTAs; "leader" = cluster rank 0):
//   full[s]   leader only. The leader's producer arms it with the bytes of BOTH CTAs; each CTA's TMA loads credit the
//             leader's barrier (cp.async.bulk.tensor...cta_group::2 with the barrier mapped into the leader).
//   empty[s]  per CTA. The leader's MMA thread releases the stage in both CTAs with one multicast tcgen05.commit.
//   tfull[a]  per CTA, same multicast commit after the last k-block of a tile.
//   tempty[a] leader only, 256 arrivals: the 128 epilogue threads of each CTA (the peer's arrive remotely).
// TMEM is allocated / released with the cta_group::2 forms by warp 1 of both CTAs; a cluster barrier precedes the release.
// Every wait is bounded (fnd_common.cuh: mbar_wait), so a mis-programmed pipeline ends with an error code, not a hang.
//
// No counterpart in the reference; the call sites are nn.Linear applications, e.g.
// src/models/fusion/cross_modal_transformer.py:147-150.
#pragma once
#include "fnd_seq_gemm.cuh"

namespace fnd {

constexpr int kSeqGemm2BN = 256;                                   // tile columns (the pair's UMMA N)
constexpr int kSeqGemm2BoxBytes = kSeqGemmBM * kSeqGemmBK * 2;       // one 128-row x 64-deep operand box: 16 KB
constexpr int kSeqGemm2StageBytes = 2 * kSeqGemm2BoxBytes;           // A box + half-W box per 64-deep k-box: 32 KB per CTA

// kKB = 64-deep k-boxes per pipeline stage (1: six 32 KB stages, five beside the residual slabs; 2: three 64 KB stages with
// 8 UMMAs per barrier round, FND_SEQ_GEMM_KB=2 — measured no faster: 161.5 vs 158.4 us). kDbg compiles the probe aids in
// (clock64 totals, work-skipping flags): the issue loop is latency-sensitive enough that their predicates cost 10 %.
template <int kKB, bool kDbg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSeqGemmThreads, 1) seq_gemm2_kernel(const __grid_constant__ SeqGemmParams P) {
  constexpr int kStage = kKB * kSeqGemm2StageBytes;
  // probe instantiation only (tools/gemm2_stamps.py): the issue loop is latency-bound, every extra instruction in it shows
  long long* const dbgp = kDbg ? P.dbg : nullptr;
  const int dflags = kDbg ? P.dbg_mmas : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kSeqGemmMaxStages;
  uint64_t* tfull_bar = empty_bar + kSeqGemmMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + 2);
  uint8_t* ring = smem + kSeqGemmHeader;
  uint8_t* stage = ring + P.nstages * kStage;
  uint8_t* rbuf = stage + 2 * kSeqGemmStageOutBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  constexpr int bn = kSeqGemm2BN;
  const int nstages = P.nstages, kblocks = P.kblocks;
  const int ntiles = P.tiles_m * P.tiles_n;                        // tiles_m counts 256-row super-tiles here
  constexpr uint32_t tmem_cols = 2 * bn;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    tma_prefetch_desc(&P.tmC);
    if (P.resid) tma_prefetch_desc(&P.tmR);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 256); mbar_init(&rfull_bar[a], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, tmem_cols);
    tmem_relinquish_pair();
  }
  tc_fence_before_sync();
  __syncwarp();
  cluster_sync_all();                              // both CTAs' barriers are initialised before anything is signalled remotely
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own A rows + own half of the W tile, bytes credited to the leader =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1u;
      bool ok = true;
      long long t_empty = 0;
      int n_issued = 0;
#pragma unroll 1
      for (int tile = cluster_id; tile < ntiles && ok; tile += nclusters) {
        const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
        const int am = tm * (2 * kSeqGemmBM) + static_cast<int>(rank) * kSeqGemmBM;
        const int bnr = tn * bn + static_cast<int>(rank) * (bn / 2);
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
          const long long c0 = dbgp ? clock64() : 0;
          if (!mbar_test_wait(&empty_bar[s], ph)) ok = mbar_wait_fast(&empty_bar[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
          if (dbgp) t_empty += clock64() - c0;
          if (!ok) break;
          if ((dflags & 4) && n_issued >= nstages) {       // timing probe: operands stay whatever the ring holds
            if (rank == 0) mbar_arrive(&full_bar[s]);
            if (++s == nstages) { s = 0; ph ^= 1u; }
            continue;
          }
          ++n_issued;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2u * kStage);
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
          uint8_t* sA = ring + s * kStage;
          uint8_t* sB = sA + kKB * kSeqGemm2BoxBytes;
#pragma unroll
          for (int j = 0; j < kKB; ++j) {          // a k-box past K is zero-filled by TMA (and still counts its bytes)
            tma_load_2d_pair(sA + j * kSeqGemm2BoxBytes, &P.tmA, fb, (kb * kKB + j) * kSeqGemmBK, am, kEvictNormal);
            tma_load_2d_pair(sB + j * kSeqGemm2BoxBytes, &P.tmB, fb, (kb * kKB + j) * kSeqGemmBK, bnr, kEvictLast);
          }
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
      if (dbgp && rank == 0) dbgp[cluster_id * 16 + 4] = t_empty;
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the leader CTA's warp 1 drives both SMs =================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * kSeqGemmBM, bn, 0, 0);
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(ring), 16);
      const uint32_t box_step = static_cast<uint32_t>(kSeqGemm2BoxBytes) >> 4;
      const uint32_t b_off = kKB * box_step;
      const uint32_t stage_step = static_cast<uint32_t>(kStage) >> 4;
      int s = 0, lt = 0;
      uint32_t ph = 0u;
      bool ok = true;
      long long t_tempty = 0, t_full = 0, t_issue = 0, t_fence = 0, t_sync = 0, n_it = 0;
      const long long c_start = dbgp ? clock64() : 0;
#pragma unroll 1
      for (int tile = cluster_id; tile < ntiles && ok; tile += nclusters, ++lt) {
        const int ab = lt & 1;
        const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
        const long long c0 = dbgp ? clock64() : 0;
        if (!mbar_test_wait(&tempty_bar[ab], aph ^ 1u)) ok = mbar_wait_fast(&tempty_bar[ab], aph ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (dbgp) t_tempty += clock64() - c0;
        if (!ok) break;
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(ab * bn);
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
          const long long c1 = dbgp ? clock64() : 0;
          if (!mbar_test_wait(&full_bar[s], ph)) ok = mbar_wait_fast(&full_bar[s], ph, P.err, FND_DEV_TIMEOUT_MMA);
          const long long c2 = dbgp ? clock64() : 0;
          if (!ok) break;
          tc_fence_after_sync();
          const uint32_t al = a_lo0 + static_cast<uint32_t>(s) * stage_step;
          const uint32_t bl = al + b_off;
          const long long c3 = dbgp ? clock64() : 0;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kKB; ++j) {
              const uint32_t aj = al + j * box_step, bj = bl + j * box_step;
              umma_f16_pair(tacc, desc64(aj, dhi), desc64(bj, dhi), idesc, (kb != 0 || j != 0) ? 1u : 0u);
              if (!(dflags & 1)) {
                umma_f16_pair(tacc, desc64(aj + 2u, dhi), desc64(bj + 2u, dhi), idesc, 1u);
                umma_f16_pair(tacc, desc64(aj + 4u, dhi), desc64(bj + 4u, dhi), idesc, 1u);
                umma_f16_pair(tacc, desc64(aj + 6u, dhi), desc64(bj + 6u, dhi), idesc, 1u);
              }                                    // (dbg bit 0: timing probe only — wrong results)
            }
            if (!(dflags & 8)) {
              umma_commit_pair_mc(&empty_bar[s], 3);
              if (kb == kblocks - 1) umma_commit_pair_mc(&tfull_bar[ab], 3);
            }
          }
          const long long c4 = dbgp ? clock64() : 0;
          __syncwarp();
          if (dbgp) { const long long c5 = clock64(); t_full += c2 - c1; t_issue += c4 - c3; t_fence += c3 - c2; t_sync += c5 - c4; ++n_it; }
          if (++s == nstages) { s = 0; ph ^= 1u; }
        }
      }
      if (dbgp && lane == 0) {
        long long* d = dbgp + cluster_id * 16;
        d[0] = clock64() - c_start; d[1] = t_tempty; d[2] = t_full; d[3] = t_issue; d[5] = n_it; d[12] = t_fence; d[13] = t_sync;
      }
    }
  } else {
    // ================= epilogue (both CTAs): warps 2..5, TMEM lane quarter = warp % 4 =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), 0), tempty_leader1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
    int lt = 0;
    uint32_t slab_ctr = 0;
    int pf_tile = cluster_id, pf_c = 0;
    uint32_t pf_n = 0;
    auto pf_issue = [&]() {                                 // residual slabs, two ahead (see fnd_seq_gemm.cuh)
      if (pf_tile >= ntiles) return;
      const int ptm = pf_tile / P.tiles_n, ptn = pf_tile - ptm * P.tiles_n;
      const uint32_t b = pf_n & 1u;
      mbar_arrive_expect_tx(&rfull_bar[b], kSeqGemmStageOutBytes);
      tma_load_2d(rbuf + b * kSeqGemmStageOutBytes, &P.tmR, &rfull_bar[b], ptn * bn + pf_c,
                  ptm * (2 * kSeqGemmBM) + static_cast<int>(rank) * kSeqGemmBM, kEvictFirst);
      ++pf_n;
      pf_c += 64;
      if (pf_c >= bn || ptn * bn + pf_c >= P.N) { pf_c = 0; pf_tile += nclusters; }
    };
    if (P.resid && epi_tid == 0) { pf_issue(); pf_issue(); }
    long long t_tfull = 0, t_p[4] = {0, 0, 0, 0};
    const long long e_start = dbgp ? clock64() : 0;
    for (int tile = cluster_id; tile < ntiles; tile += nclusters, ++lt) {
      const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
      const int ab = lt & 1;
      const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
      const long long c0 = dbgp ? clock64() : 0;
      const bool ok = mbar_wait(&tfull_bar[ab], aph, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      if (dbgp) t_tfull += clock64() - c0;
      tc_fence_after_sync();
      const int m0 = tm * (2 * kSeqGemmBM) + static_cast<int>(rank) * kSeqGemmBM;
      const uint32_t taddr = tmem_base + lane_addr + static_cast<uint32_t>(ab * bn);
      const int nb = tn * bn;
#pragma unroll 1
      for (int c = 0; c < bn; c += 64) {
        const int n0 = nb + c;
        if (n0 >= P.N || (dflags & 2)) break;          // tile-uniform (dbg bit 1: timing probe without the epilogue's work)
        const uint32_t sb = slab_ctr & 1u;
        uint8_t* stg = stage + sb * kSeqGemmStageOutBytes;
        const long long p0 = dbgp ? clock64() : 0;
        if (epi_tid == 0) tma_store_wait_read<1>();        // the store issued two slabs ago has drained this buffer
        epi_bar_sync();
        if (P.resid) mbar_wait(&rfull_bar[sb], (slab_ctr >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        const long long p1 = dbgp ? clock64() : 0;
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
        const long long p2 = dbgp ? clock64() : 0;
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
        const long long p3 = dbgp ? clock64() : 0;
        fence_proxy_async_smem();
        epi_bar_sync();
        if (epi_tid == 0) {
          if (ok) {
            tma_store_2d(&P.tmC, stg, n0, m0);             // rows >= M and columns >= N are clipped by the tensor map
            tma_store_commit();
          }
          if (P.resid) pf_issue();
        }
        if (dbgp) { t_p[0] += p1 - p0; t_p[1] += p2 - p1; t_p[2] += p3 - p2; t_p[3] += clock64() - p3; }
        ++slab_ctr;
      }
      // this CTA's half of the accumulator is in registers / staged: tell the leader's MMA warp
      tc_fence_before_sync();
      mbar_arrive_cluster(ab ? tempty_leader1 : tempty_leader0);
    }
    if (dbgp && epi_tid == 0 && rank == 0) { dbgp[cluster_id * 16 + 6] = clock64() - e_start; dbgp[cluster_id * 16 + 7] = t_tfull;
      for (int i = 0; i < 4; ++i) dbgp[cluster_id * 16 + 8 + i] = t_p[i]; }
    if (epi_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before_sync();
  __syncwarp();
  cluster_sync_all();                              // neither CTA may release the pair's TMEM while the other still reads it
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

}  // namespace fnd
