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
//     epilogue warps drain tile i (tcgen05.ld -> bias / residual / activation -> packed bf16 256-bit stores),
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
constexpr int kSeqGemmRingBudget = 200 * 1024;

struct alignas(64) SeqGemmParams {
  CUtensorMap tmA, tmB;
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
};

__global__ void __launch_bounds__(kSeqGemmThreads, 1) seq_gemm_kernel(const __grid_constant__ SeqGemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kSeqGemmMaxStages;
  uint64_t* tfull_bar = empty_bar + kSeqGemmMaxStages;     // [2] accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* ring = smem + kSeqGemmHeader;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = P.bn, nstages = P.nstages, stage_bytes = P.stage_bytes, kblocks = P.kblocks;
  const int ntiles = P.tiles_m * P.tiles_n;
  const uint32_t tmem_cols = static_cast<uint32_t>(2 * bn);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
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
      int it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
        const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % nstages;
          const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
          ok = mbar_wait(&empty_bar[s], ph ^ 1u, P.err, FND_DEV_TIMEOUT_PRODUCER);
          if (!ok) break;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          uint8_t* sA = ring + s * stage_bytes;
          tma_load_2d(sA, &P.tmA, &full_bar[s], kb * kSeqGemmBK, tm * kSeqGemmBM, kEvictNormal);
          tma_load_2d(sA + kSeqGemmBM * kSeqGemmBK * 2, &P.tmB, &full_bar[s], kb * kSeqGemmBK, tn * bn, kEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kSeqGemmBM, bn, 0, 0);
      int it = 0, lt = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++lt) {
        const int ab = lt & 1;
        const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
        ok = mbar_wait(&tempty_bar[ab], aph ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(ab * bn);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % nstages;
          const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
          ok = mbar_wait(&full_bar[s], ph, P.err, FND_DEV_TIMEOUT_MMA);
          if (!ok) break;
          tc_fence_after_sync();
          const uint32_t aBase = smem_u32(ring + s * stage_bytes);
          const uint32_t bBase = aBase + kSeqGemmBM * kSeqGemmBK * 2;
#pragma unroll
          for (int k = 0; k < kSeqGemmBK / 16; ++k) {
            const uint64_t ad = make_smem_desc_sw128(aBase + k * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(bBase + k * 32, 16, 1024);
            umma_f16(tacc, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[ab]);
      }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quarter = warp % 4 =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
      const int ab = lt & 1;
      const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
      const bool ok = mbar_wait(&tfull_bar[ab], aph, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      const int m = tm * kSeqGemmBM + row;
      const bool row_ok = ok && m < P.M;
      const uint32_t taddr = tmem_base + lane_addr + static_cast<uint32_t>(ab * bn);
      const int nb = tn * bn;
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
        if (P.out_bf) {
          __nv_bfloat16* op = P.out_bf + static_cast<size_t>(m) * P.out_pitch + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols)
              *reinterpret_cast<uint4*>(op + j) = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                                             pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
          }
        }
        if (P.out_f32) {
          float* op = P.out_f32 + static_cast<size_t>(m) * P.f32_pitch + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      // every column of this accumulator is in registers / stored: hand it back to the MMA warp
      tc_fence_before_sync();
      mbar_arrive(&tempty_bar[ab]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace fnd
