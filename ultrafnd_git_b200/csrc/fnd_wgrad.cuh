// fnd_wgrad.cuh — persistent weight-gradient kernel (the "light" GEMM launches: dW = dY^T X, both operands MN-major,
// contraction = batch <= 192, no split-K).
//
// At the reference's batch a wgrad tile is only 2 k-blocks of MMA followed by a 64 KB fp32 store: launched one CTA
// per tile (780 of them) the fixed per-CTA costs — barrier init, TMEM alloc/dealloc, CTA scheduling, the serial
// load -> MMA -> store chain — dominated (31 us for 51 MB of output). Here at most 2 x 148 CTAs stay resident and walk
// the tile list (tile = blockIdx, blockIdx + grid, ...), with the three roles decoupled across tiles:
//   * TMA warp   keeps a 2-stage operand ring full, running ahead into the next tile;
//   * MMA warp   alternates between TWO TMEM accumulators (2 x 128 columns), so tile i+1 is being multiplied while
//   * the eight epilogue warps drain tile i: each warp transposes 32x32 blocks through its own staging buffer and
//     writes 4 full rows x 128 B per store instruction, accumulates the sum of squares for the gradient norm, and hands
//     the accumulator back through an mbarrier.
// The trailing CTAs [gemm_ctas, gridDim.x) run the finalize jobs (fnd_rows.cuh) exactly as before.
//
// Replaces autograd's weight-gradient addmm calls of the reference step (src/training/forensic_trainer.py:291).
#pragma once
#include "fnd_gemm.cuh"

namespace fnd {

constexpr int kWgStages = 2;
constexpr int kWgStageBytes = kGemmStageBytesA + 128 * kGemmBK * 2;            // 32 KB (bn = 128)
constexpr int kWgStgPitch = 36;                                                // floats per staged row
constexpr int kWgStagingBytes = kGemmEpiWarps * 32 * kWgStgPitch * 4;          // 36 KB
constexpr int kWgSmemBytes = 1024 /*align slack*/ + kGemmSmemHeader + kWgStages * kWgStageBytes + kWgStagingBytes;
constexpr int kWgTmemCols = 256;
constexpr int kWgMaxCtas = 2 * 148;

__global__ void __launch_bounds__(kGemmThreads, 2)
fnd_wgrad_kernel(const __grid_constant__ GemmTableP tbl, RunCtx ctx, const __grid_constant__ FinParams fin, int total_tiles) {
  if (static_cast<int>(blockIdx.x) >= tbl.gemm_ctas) {
    griddep_wait();
    griddep_launch();
    if (threadIdx.x < 256) finalize_cta(fin, static_cast<int>(blockIdx.x) - tbl.gemm_ctas, static_cast<int>(gridDim.x) - tbl.gemm_ctas);
    return;
  }
  const GemmProblem* probs = tbl.p;
  const int nprob = tbl.nprob;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);        // [2]
  uint64_t* empty_bar = full_bar + kWgStages;                    // [2]
  uint64_t* acc_full = empty_bar + kWgStages;                    // [2]
  uint64_t* acc_empty = acc_full + 2;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* red_smem = reinterpret_cast<float*>(tmem_slot + 2);     // [2][8]
  uint8_t* ring = smem + kGemmSmemHeader;
  float* staging = reinterpret_cast<float*>(ring + kWgStages * kWgStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nctas = tbl.gemm_ctas;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nprob; ++i) {
      tma_prefetch_desc(&probs[i].tmA[0]);
      tma_prefetch_desc(&probs[i].tmB[0]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kWgStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc_full[b], 1);
        mbar_init(&acc_empty[b], kGemmEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kWgTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto locate = [&](int t, int& pi, int& tm, int& tn) {
    pi = 0;
    while (pi + 1 < nprob && t >= probs[pi + 1].cta_begin) ++pi;
    const int local = t - probs[pi].cta_begin;
    tm = local % probs[pi].tiles_m;
    tn = local / probs[pi].tiles_m;
  };

  if (warp == 0) {
    // ================= TMA producer (runs ahead across tiles) =================
    if (lane == 0) {
      griddep_wait();
      griddep_launch();
      uint32_t it = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < total_tiles && ok; t += nctas) {
        int pi, tm, tn;
        locate(t, pi, tm, tn);
        const GemmProblem& P = probs[pi];
        const int bn = P.bn, ncombo = P.ncombo;
        const uint32_t tx = kGemmStageBytesA + static_cast<uint32_t>(bn) * kGemmBK * 2;
        const int iters = P.kb_total * ncombo;
        for (int j = 0; j < iters; ++j, ++it) {
          const int s = it % kWgStages;
          const uint32_t ph = (it / kWgStages) & 1u;
          ok = mbar_wait(&empty_bar[s], ph ^ 1u, ctx.err, FND_DEV_TIMEOUT_PRODUCER);
          if (!ok) break;
          const int kb = j / ncombo, c = j - kb * ncombo;
          const void* mapA = &P.tmA[c == 2 ? 1 : 0];
          const void* mapB = &P.tmB[c == 1 ? 1 : 0];
          uint8_t* sA = ring + s * kWgStageBytes;
          uint8_t* sB = sA + kGemmStageBytesA;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          tma_load_2d(sA, mapA, &full_bar[s], tm * kGemmBM, kb * kGemmBK, P.hintA);
          tma_load_2d(sA + 8192, mapA, &full_bar[s], tm * kGemmBM + 64, kb * kGemmBK, P.hintA);
          for (int ch = 0; ch < bn / 64; ++ch)
            tma_load_2d(sB + ch * 8192, mapB, &full_bar[s], tn * bn + ch * 64, kb * kGemmBK, P.hintB);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (two accumulators) =================
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles && ok; t += nctas, ++i) {
        int pi, tm, tn;
        locate(t, pi, tm, tn);
        const GemmProblem& P = probs[pi];
        const uint32_t idesc = make_idesc_bf16(kGemmBM, P.bn, 1, 1);
        const int iters = P.kb_total * P.ncombo;
        const int b = i & 1;
        ok = mbar_wait(&acc_empty[b], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u, ctx.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * 128);
        for (int j = 0; j < iters; ++j, ++it) {
          const int s = it % kWgStages;
          const uint32_t ph = (it / kWgStages) & 1u;
          ok = mbar_wait(&full_bar[s], ph, ctx.err, FND_DEV_TIMEOUT_MMA);
          if (!ok) break;
          tc_fence_after_sync();
          const uint32_t aBase = smem_u32(ring + s * kWgStageBytes);
          const uint32_t bBase = aBase + kGemmStageBytesA;
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // MN-major operands: 16 contraction rows of 128 B = 2048 B (two 8-row swizzle atoms)
            const uint64_t ad = make_smem_desc_sw128(aBase + k * 2048u, 8192u, 1024);
            const uint64_t bd = make_smem_desc_sw128(bBase + k * 2048u, 8192u, 1024);
            umma_f16(tacc, ad, bd, idesc, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..9) =================
    const int ew = warp - 2;
    const int lane_grp = warp & 3;
    const int half = ew >> 2;
    const int epi_tid = threadIdx.x - 64;
    float* stg = staging + ew * (32 * kWgStgPitch);
    const int rsub = lane >> 3, csub = (lane & 7) * 4;
    griddep_wait();
    EpiCtx X;                                   // wgrad epilogues never use dropout
    X.dfw = make_dropcfg(0.f, 0ull); X.dbw = X.dfw; X.key_fw = 0u; X.key_bw = 0u;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += nctas, ++i) {
      int pi, tm, tn;
      locate(t, pi, tm, tn);
      const GemmProblem& P = probs[pi];
      const EpiParams E = P.epi;
      const int PM = P.M, PN = P.N, bn = P.bn;
      const int local = t - P.cta_begin;
      const int b = i & 1;
      const int row = lane_grp * 32 + lane;
      const int m = tm * kGemmBM + row;
      const bool row_ok = m < PM;
      const int nb = tn * bn;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(b * 128);
      const bool proceed = mbar_wait(&acc_full[b], static_cast<uint32_t>(i >> 1) & 1u, ctx.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      float ss = 0.f;
      if (E.plain_f32 && bn >= 64) {
        const int c_end = (half + 1) * (bn >> 1);
#pragma unroll 1
        for (int c = half * (bn >> 1); c < c_end; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v4 = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                          __uint_as_float(r[4 * q + 3]));
            *reinterpret_cast<float4*>(stg + lane * kWgStgPitch + 4 * q) = v4;
            if (proceed && row_ok && nb + c + 4 * q < PN)
              ss = fmaf(v4.x, v4.x, fmaf(v4.y, v4.y, fmaf(v4.z, v4.z, fmaf(v4.w, v4.w, ss))));
          }
          __syncwarp();
          const int n = nb + c + csub;
          if (proceed && n < PN) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = k * 4 + rsub;
              const int mm = tm * kGemmBM + lane_grp * 32 + rr;
              if (mm < PM)
                *reinterpret_cast<float4*>(E.out_f32 + static_cast<size_t>(mm) * E.f32_pitch + n) =
                    *reinterpret_cast<const float4*>(stg + rr * kWgStgPitch + csub);
            }
          }
          __syncwarp();
        }
      } else {
        const int ngroups = bn / 8;
#pragma unroll 1
        for (int g = half; g < ngroups; g += 2) {
          uint32_t r[8];
          tmem_ld_32x8(taddr + g * 8, r);
          tmem_ld_wait();
          const int n0 = nb + g * 8;
          if (!proceed || !row_ok || n0 >= PN) continue;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
          ss += epi_group(E, X, v, m, n0, PN, 0.f, 0.f);
        }
      }
      // this warp is done reading accumulator b: hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
      if (E.done_ctr) __threadfence();
      if (E.sumsq_slots || E.done_ctr) {
        ss = warp_sum(ss);
        if (lane == 0) red_smem[b * 8 + ew] = ss;
        epi_named_barrier();
        if (epi_tid == 0) {
          const float* rs = red_smem + b * 8;
          if (E.sumsq_slots) E.sumsq_slots[local] = ((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + (rs[6] + rs[7]));
          if (E.done_ctr) atomicAdd(E.done_ctr, 1u);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

}  // namespace fnd
