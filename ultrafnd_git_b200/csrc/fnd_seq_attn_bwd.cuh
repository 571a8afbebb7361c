// fnd_seq_attn_bwd.cuh — fused backward of the multi-head cross-attention (Tier B), two tcgen05 kernels that share the
// forward kernel's skeleton (fnd_seq_attn.cuh: persistent CTA, two 128-row tiles per work item each with its own softmax
// warpgroup and TMEM regions, TMA ring, one MMA-issuing thread, A operands of the second GEMMs read from TENSOR MEMORY):
//
//   P  = exp2(S * scale_log2 - lse2)            recomputed from the forward's logsumexp, never stored
//   dP = dO V^T,   dS = P o (dP - D),           D = rowsum(dO o O)   (seq_attn_bwd_prep_kernel)
//   dQ = scale * dS K         seq_attn_bwd_dq_kernel : item = (sample, head, 256 QUERY rows), walks the key blocks of 64;
//                             per tile and block  S = Q K_j^T,  dP = dO V_j^T  (TMEM) -> one thread per query row forms
//                             dS as bf16 pairs in place of its S row -> dQ += dS K_j (A from TMEM, K_j MN-major)
//   dK = scale * dS^T Q       seq_attn_bwd_dkv_kernel: item = (sample, head, 256 KEY rows), walks the query blocks of 64;
//   dV = P^T dO               per tile and block  S^T = K Q_i^T,  dP^T = V dO_i^T -> one thread per KEY row (the
//                             per-query lse2 / D come from the ring stage in shared memory) writes P^T over its S^T row and
//                             dS^T over its dP^T row -> dV += P^T dO_i,  dK += dS^T Q_i   (A from TMEM, B MN-major)
// Seven GEMMs instead of the five of a single-pass backward, but dQ needs no atomics / no fp32 global accumulation and
// every result is deterministic. Masked keys and padded queries contribute exactly 0 (score -inf, lse2 = +inf).
//
// No counterpart in the reference; checked against torch autograd over the self-oracle oracle/seq_oracle.py.
#pragma once
#include "fnd_seq_attn.cuh"

namespace fnd {

constexpr int kBwdBlk = 64;                              // rows of the streamed operand per block (keys for dQ, queries for dK/dV)
constexpr int kBwdStages = 4;
constexpr int kBwdTileBytes = kAttnBQ * kAttnD * 2;      // 16 KB: a resident 128 x 64 tile
constexpr int kBwdBlkBytes = kBwdBlk * kAttnD * 2;       // 8 KB: a streamed 64 x 64 tile
constexpr int kBwdStageBytes = 2 * kBwdBlkBytes + 1024;  // two streamed tiles + 64 lse2 + 64 D values (dK/dV kernel); 1024-aligned for SWIZZLE_128B
constexpr int kBwdSmemBytes = 1024 + 1024 + 2 * 4 * kBwdTileBytes + kBwdStages * kBwdStageBytes;
constexpr int kAttnBwdDefaultPingPong = 1;              // FND_ATTN_BWD_PINGPONG overrides
constexpr int kBwdTmemS = 0, kBwdTmemDP = 128, kBwdTmemAcc0 = 256, kBwdTmemAcc1 = 384;

struct alignas(64) AttnBwdParams {
  // dQ kernel : tmR0 = Q, tmR1 = dO (resident, box 128 rows); tmS0 = K, tmS1 = V (streamed, box 64 rows)
  // dKV kernel: tmR0 = K, tmR1 = V (resident);                tmS0 = Q, tmS1 = dO (streamed)
  CUtensorMap tmR0, tmR1, tmS0, tmS1;
  int B, H, Lq, Lk, Lqp;
  int r0_col0, r1_col0, s0_col0, s1_col0;    // first column of head 0 inside each matrix
  const int* kv_len;
  const unsigned char* kv_mask;
  float scale_log2, scale;
  const float* lse2p;                        // [B, H, Lqp] logsumexp in exp2 units, +inf = "row contributes nothing"
  const float* Dp;                           // [B, H, Lqp]
  __nv_bfloat16* out0; int out0_pitch; int out0_col0;   // dQ | dK
  __nv_bfloat16* out1; int out1_pitch; int out1_col0;   // -  | dV
  int* err;
  int pingpong;                              // 1: the two warpgroups take turns in the exp2 section (as in the forward kernel)
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kDkv = false: dQ kernel; true: dK / dV kernel. The two differ in what is resident / streamed, in where the per-row and
// per-column softmax statistics come from, and in the number of accumulators; the pipeline is the same.
template <bool kDkv>
__global__ void __launch_bounds__(kAttnThreads, 1) seq_attn_bwd_kernel(const __grid_constant__ AttnBwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* r_full = reinterpret_cast<uint64_t*>(smem);      // [2]: resident tiles of an item, double-buffered across items
  uint64_t* r_empty = r_full + 2;
  uint64_t* st_full = r_empty + 2;                           // [stages]
  uint64_t* st_empty = st_full + kBwdStages;
  uint64_t* s_full = st_empty + kBwdStages;                  // [2]: per tile
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  uint8_t* sR = smem + 1024;                                 // [2 buffers][R0 tile0, R0 tile1, R1 tile0, R1 tile1]
  uint8_t* sS = sR + 2 * 4 * kBwdTileBytes;                  // [stages][S0 | S1 | lse2[64] | D[64]]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lres = kDkv ? P.Lk : P.Lq;                       // rows of the resident side
  const int nrt = (Lres + kAttnItemQ - 1) / kAttnItemQ;
  const int nwork = nrt * P.H * P.B;
  // Blocks of the streamed side an item walks; 0 = the item produces zeros without touching the tensor pipe.
  auto item_nblk = [&](int w) -> int {
    const int b = w / (nrt * P.H);
    int kv_len = P.Lk;
    if (P.kv_len) kv_len = min(max(P.kv_len[b], 0), P.Lk);
    if (kDkv) {
      const int rt = w % nrt;
      if (rt * kAttnItemQ >= kv_len) return 0;               // every key of this item is padding
      return (P.Lq + kBwdBlk - 1) / kBwdBlk;
    }
    return (kv_len + kBwdBlk - 1) / kBwdBlk;
  };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&P.tmR0); tma_prefetch_desc(&P.tmR1);
    tma_prefetch_desc(&P.tmS0); tma_prefetch_desc(&P.tmS1);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 2); }      // "empty": one commit per MMA warp
      for (int s = 0; s < kBwdStages; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 2); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kAttnTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<72>();
    if (warp == 8) {
      // ================= TMA producer =================
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 1u;
        uint32_t rn = 0;
        bool ok = true;
#pragma unroll 1
        for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
          const int nblk = item_nblk(w);
          if (nblk == 0) continue;
          const int rt = w % nrt, h = (w / nrt) % P.H, b = w / (nrt * P.H);
          ok = mbar_wait_fast(&r_empty[rn & 1u], ((rn >> 1) & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_PRODUCER);
          if (!ok) break;
          mbar_arrive_expect_tx(&r_full[rn & 1u], 4 * kBwdTileBytes);
          uint8_t* r = sR + (rn & 1u) * 4 * kBwdTileBytes;
          const int row0 = rt * kAttnItemQ;
          tma_load_3d(r, &P.tmR0, &r_full[rn & 1u], P.r0_col0 + h * kAttnD, row0, b, kEvictFirst);
          tma_load_3d(r + kBwdTileBytes, &P.tmR0, &r_full[rn & 1u], P.r0_col0 + h * kAttnD, row0 + kAttnBQ, b, kEvictFirst);
          tma_load_3d(r + 2 * kBwdTileBytes, &P.tmR1, &r_full[rn & 1u], P.r1_col0 + h * kAttnD, row0, b, kEvictFirst);
          tma_load_3d(r + 3 * kBwdTileBytes, &P.tmR1, &r_full[rn & 1u], P.r1_col0 + h * kAttnD, row0 + kAttnBQ, b, kEvictFirst);
          ++rn;
          const float* lrow = P.lse2p + (static_cast<size_t>(b) * P.H + h) * P.Lqp;
          const float* drow = P.Dp + (static_cast<size_t>(b) * P.H + h) * P.Lqp;
#pragma unroll 1
          for (int j = 0; j < nblk; ++j) {
            uint8_t* st = sS + s * kBwdStageBytes;
            ok = mbar_wait_fast(&st_empty[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
            if (!ok) break;
            mbar_arrive_expect_tx(&st_full[s], 2 * kBwdBlkBytes + (kDkv ? 512 : 0));
            tma_load_3d(st, &P.tmS0, &st_full[s], P.s0_col0 + h * kAttnD, j * kBwdBlk, b, kEvictLast);
            tma_load_3d(st + kBwdBlkBytes, &P.tmS1, &st_full[s], P.s1_col0 + h * kAttnD, j * kBwdBlk, b, kEvictLast);
            if (kDkv) {
              bulk_load_1d(st + 2 * kBwdBlkBytes, lrow + j * kBwdBlk, 256, &st_full[s]);
              bulk_load_1d(st + 2 * kBwdBlkBytes + 256, drow + j * kBwdBlk, 256, &st_full[s]);
            }
            if (++s == kBwdStages) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ================= MMA issuers: one warp PER TILE (warp 9: tile 0, warp 10: tile 1), as in the forward kernel =================
      // Each warp follows its own tile:  S_t, dP_t of the first block;  then per block  p_full_t -> accumulating GEMMs ->
      // S_t, dP_t of the next block. The operand waits of the next block are done BEFORE the critical wait on p_full (an
      // mbarrier try_wait costs ~90 cycles even on a completed phase). Shared buffers (resident tiles, ring stages) are
      // released by one tcgen05.commit arrival from each warp.
      const int t = warp - 9;
      const uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kBwdBlk, 0, 0);      // S / dP: 128 x 64 x 64, both operands K-major
      const uint32_t idesc_a = make_idesc_bf16(kAttnBQ, kAttnD, 0, 1);       // accumulators: A from TMEM, B MN-major [rows][d]
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      const uint32_t r_lo = smem_desc_lo(smem_u32(sR), 16) + static_cast<uint32_t>(t) * (kBwdTileBytes >> 4);
      const uint32_t s_lo = smem_desc_lo(smem_u32(sS), 16);                  // K-major view of a streamed tile
      const uint32_t s_mn = smem_desc_lo(smem_u32(sS), 8192);                // MN-major view of the same tile
      const uint32_t tS = tmem_base + kBwdTmemS + static_cast<uint32_t>(t) * kBwdBlk;
      const uint32_t tD = tmem_base + kBwdTmemDP + static_cast<uint32_t>(t) * kBwdBlk;
      const uint32_t tA0 = tmem_base + kBwdTmemAcc0 + static_cast<uint32_t>(t) * kAttnD;
      const uint32_t tA1 = tmem_base + kBwdTmemAcc1 + static_cast<uint32_t>(t) * kAttnD;
      bool ok = true;
      int ws = blockIdx.x, js = 0, ns = 0;
      int ks = 0; uint32_t kph = 0u;
      uint32_t rn = 0, rcur = 0;
      bool s_ready = false;
      auto s_next_item = [&]() {
        for (; ws < nwork; ws += gridDim.x) {
          ns = item_nblk(ws);
          if (ns > 0) return;
        }
        ns = 0;
      };
      auto prepare_s = [&]() {                                               // operand waits of the next S / dP step
        if (ns == 0 || s_ready) return;
        if (js == 0) {
          rcur = rn & 1u;
          ok = ok && mbar_wait_fast(&r_full[rcur], (rn >> 1) & 1u, P.err, FND_DEV_TIMEOUT_MMA);
          ++rn;
        }
        ok = ok && mbar_wait_fast(&st_full[ks], kph, P.err, FND_DEV_TIMEOUT_MMA);
        s_ready = true;
      };
      // S_t = R0_t S0^T and dP_t = R1_t S1^T of block (ws, js)
      auto issue_s = [&]() {
        if (ns == 0) return;
        prepare_s();
        tc_fence_after_sync();
        const uint32_t a0 = r_lo + rcur * 4u * (kBwdTileBytes >> 4);
        const uint32_t a1 = a0 + 2u * (kBwdTileBytes >> 4);
        const uint32_t b0 = s_lo + static_cast<uint32_t>(ks) * (kBwdStageBytes >> 4);
        const uint32_t b1 = b0 + (kBwdBlkBytes >> 4);
        if (ok && elect_one()) {
#pragma unroll
          for (int k = 0; k < kAttnD / 16; ++k) umma_f16(tS, desc64(a0 + 2 * k, dhi), desc64(b0 + 2 * k, dhi), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kAttnD / 16; ++k) umma_f16(tD, desc64(a1 + 2 * k, dhi), desc64(b1 + 2 * k, dhi), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          if (js + 1 == ns) umma_commit(&r_empty[rcur]);                     // second arrival: the other tile's warp
        }
        __syncwarp();
        if (++ks == kBwdStages) { ks = 0; kph ^= 1u; }
        if (++js == ns) { js = 0; ws += gridDim.x; s_next_item(); }
        s_ready = false;
      };
      s_next_item();
      issue_s();
      uint32_t gp = 0, ip = 0;
      int vs = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
        const int nblk = item_nblk(w);
        if (nblk == 0) continue;
#pragma unroll 1
        for (int j = 0; j < nblk && ok; ++j) {
          prepare_s();                                                       // next block's operands: off the critical path
          if (j == 0) ok = ok && mbar_wait_fast(&o_free[t], (ip & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
          ok = ok && mbar_wait_fast(&p_full[t], gp & 1u, P.err, FND_DEV_TIMEOUT_MMA);
          if (!ok) break;
          tc_fence_after_sync();
          const uint32_t bm = s_mn + static_cast<uint32_t>(vs) * (kBwdStageBytes >> 4);
          const uint32_t acc = (j != 0) ? 1u : 0u;
          if (elect_one()) {
            if (kDkv) {
              // dV_t += P^T_t dO_i   (P^T over the S^T columns, dO_i = streamed tile 1)
              // dK_t += dS^T_t Q_i   (dS^T over the dP^T columns, Q_i = streamed tile 0)
#pragma unroll
              for (int k = 0; k < kBwdBlk / 16; ++k)
                umma_f16_ts(tA1, tS + 8 * k, desc64(bm + (kBwdBlkBytes >> 4) + 128 * k, dhi), idesc_a, (acc | (k != 0)) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < kBwdBlk / 16; ++k)
                umma_f16_ts(tA0, tD + 8 * k, desc64(bm + 128 * k, dhi), idesc_a, (acc | (k != 0)) ? 1u : 0u);
            } else {
              // dQ_t += dS_t K_j     (dS over the S columns, K_j = streamed tile 0)
#pragma unroll
              for (int k = 0; k < kBwdBlk / 16; ++k)
                umma_f16_ts(tA0, tS + 8 * k, desc64(bm + 128 * k, dhi), idesc_a, (acc | (k != 0)) ? 1u : 0u);
            }
            if (j + 1 == nblk) umma_commit(&o_full[t]);
            umma_commit(&st_empty[vs]);                                      // second arrival: the other tile's warp
          }
          __syncwarp();
          issue_s();
          ++gp;
          if (++vs == kBwdStages) vs = 0;
        }
        ++ip;
      }
    }
  } else {
    // ================= one thread per resident row: warps 0..7 =================
    setmaxnreg_inc<216>();
    const int t = warp >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + kBwdTmemS + static_cast<uint32_t>(t) * kBwdBlk;
    const uint32_t tD = tmem_base + lane_addr + kBwdTmemDP + static_cast<uint32_t>(t) * kBwdBlk;
    const uint32_t tA0 = tmem_base + lane_addr + kBwdTmemAcc0 + static_cast<uint32_t>(t) * kAttnD;
    const uint32_t tA1 = tmem_base + lane_addr + kBwdTmemAcc1 + static_cast<uint32_t>(t) * kAttnD;
    const uint64_t sl2 = pack_f32x2(P.scale_log2, P.scale_log2);
    const uint32_t sS_s = smem_u32(sS);
    bool ok = true;
    uint32_t g = 0, ip = 0;
    const bool pingpong = P.pingpong != 0;       // exp2-section token, see fnd_seq_attn.cuh
    if (pingpong && t == 1) named_bar_arrive(1, 256);
#pragma unroll 1
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int rt = w % nrt, h = (w / nrt) % P.H, b = w / (nrt * P.H);
      int kv_len = P.Lk;
      if (P.kv_len) kv_len = min(max(__ldg(P.kv_len + b), 0), P.Lk);
      const int ri = rt * kAttnItemQ + t * kAttnBQ + row;       // this thread's row of the resident side
      const unsigned char* mrow = P.kv_mask ? P.kv_mask + static_cast<size_t>(b) * P.Lk : nullptr;
      int nblk;
      float my_lse2 = INFINITY, my_D = 0.f;                     // dQ kernel: this query row's statistics
      bool my_valid = true;                                     // dK/dV kernel: this key takes part at all
      if (kDkv) {
        nblk = (rt * kAttnItemQ >= kv_len) ? 0 : (P.Lq + kBwdBlk - 1) / kBwdBlk;
        my_valid = ri < kv_len && (!mrow || mrow[ri] != 0);
      } else {
        nblk = (kv_len + kBwdBlk - 1) / kBwdBlk;
        if (ri < P.Lqp) {
          const size_t o = (static_cast<size_t>(b) * P.H + h) * P.Lqp + ri;
          my_lse2 = __ldg(P.lse2p + o);
          my_D = __ldg(P.Dp + o);
        }
      }
      const uint64_t nl2 = pack_f32x2(-my_lse2, -my_lse2);
      const uint64_t nD2 = pack_f32x2(-my_D, -my_D);

#pragma unroll 1
      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t stage = g % kBwdStages;
        if (kDkv) ok = ok && mbar_wait_fast(&st_full[stage], (g / kBwdStages) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);   // lse2 / D of this block
        ok = ok && mbar_wait_fast(&s_full[t], g & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        tc_fence_after_sync();
        float s[kBwdBlk], dp[kBwdBlk];
        tmem_ld_32x32(tS, reinterpret_cast<uint32_t(&)[32]>(s[0]));
        tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
        tmem_ld_32x32(tD, reinterpret_cast<uint32_t(&)[32]>(dp[0]));
        tmem_ld_32x32(tD + 32, reinterpret_cast<uint32_t(&)[32]>(dp[32]));
        tmem_ld_wait();
        uint32_t pk[kBwdBlk / 2], dk[kBwdBlk / 2];
        if (pingpong) named_bar_sync(1 + t, 256);
        if (kDkv) {
          // columns = the block's 64 queries: lse2 / D per column from the ring stage (broadcast 128-bit loads)
          const uint32_t stat = sS_s + stage * kBwdStageBytes + 2 * kBwdBlkBytes;
          const float vmask = my_valid ? 1.f : 0.f;
#pragma unroll
          for (int i = 0; i < kBwdBlk; i += 4) {
            const uint4 l4 = lds_v4(stat + i * 4), d4 = lds_v4(stat + 256 + i * 4);
            float x0, x1, x2, x3;
            unpack_f32x2(fma_f32x2(pack_f32x2(s[i], s[i + 1]), sl2, pack_f32x2(-__uint_as_float(l4.x), -__uint_as_float(l4.y))), x0, x1);
            unpack_f32x2(fma_f32x2(pack_f32x2(s[i + 2], s[i + 3]), sl2, pack_f32x2(-__uint_as_float(l4.z), -__uint_as_float(l4.w))), x2, x3);
            const float p0 = ex2_approx(x0) * vmask, p1 = ex2_approx(x1) * vmask, p2 = ex2_approx(x2) * vmask, p3 = ex2_approx(x3) * vmask;
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            dk[i >> 1] = pack_bf16x2(p0 * (dp[i] - __uint_as_float(d4.x)), p1 * (dp[i + 1] - __uint_as_float(d4.y)));
            dk[(i >> 1) + 1] = pack_bf16x2(p2 * (dp[i + 2] - __uint_as_float(d4.z)), p3 * (dp[i + 3] - __uint_as_float(d4.w)));
          }
          tmem_st_32x32(tS, pk);                                 // P^T over this row of S^T
          tmem_st_32x32(tD, dk);                                 // dS^T over this row of dP^T
        } else {
          // columns = the block's 64 keys: key-padding mask as two validity words
          const int k0 = j * kBwdBlk;
          uint32_t va = 0xffffffffu, vb = 0xffffffffu;
          if (mrow || k0 + kBwdBlk > kv_len) {
            const int ka = k0 + lane, kb = ka + 32;
            va = __ballot_sync(0xffffffffu, ka < kv_len && (!mrow || mrow[ka] != 0));
            vb = __ballot_sync(0xffffffffu, kb < kv_len && (!mrow || mrow[kb] != 0));
          }
          if ((va & vb) != 0xffffffffu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = ((va >> i) & 1u) ? s[i] : -INFINITY;
              s[32 + i] = ((vb >> i) & 1u) ? s[32 + i] : -INFINITY;
            }
          }
#pragma unroll
          for (int i = 0; i < kBwdBlk; i += 4) {
            float x0, x1, x2, x3, e0, e1, e2, e3;
            unpack_f32x2(fma_f32x2(pack_f32x2(s[i], s[i + 1]), sl2, nl2), x0, x1);
            unpack_f32x2(fma_f32x2(pack_f32x2(s[i + 2], s[i + 3]), sl2, nl2), x2, x3);
            unpack_f32x2(add_f32x2(pack_f32x2(dp[i], dp[i + 1]), nD2), e0, e1);
            unpack_f32x2(add_f32x2(pack_f32x2(dp[i + 2], dp[i + 3]), nD2), e2, e3);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
            dk[i >> 1] = pack_bf16x2(p0 * e0, p1 * e1);
            dk[(i >> 1) + 1] = pack_bf16x2(p2 * e2, p3 * e3);
          }
          tmem_st_32x32(tS, dk);                                 // dS over this row of S
        }
        if (pingpong) named_bar_arrive(2 - t, 256);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      // ---- item epilogue: accumulators * scale -> bf16 -> global (one 128-byte row segment per thread and output) ----
      const bool row_ok = ri < Lres;
      if (nblk > 0) {
        ok = ok && mbar_wait_fast(&o_full[t], ip & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        tc_fence_after_sync();
        uint32_t r0[kAttnD], r1[kAttnD];
        tmem_ld_32x32(tA0, reinterpret_cast<uint32_t(&)[32]>(r0[0]));
        tmem_ld_32x32(tA0 + 32, reinterpret_cast<uint32_t(&)[32]>(r0[32]));
        if (kDkv) {
          tmem_ld_32x32(tA1, reinterpret_cast<uint32_t(&)[32]>(r1[0]));
          tmem_ld_32x32(tA1 + 32, reinterpret_cast<uint32_t(&)[32]>(r1[32]));
        }
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
        ++ip;
        if (row_ok) {
          const float sc0 = ok ? P.scale : 0.f;                  // dQ and dK carry the softmax scale
          __nv_bfloat16* o0 = P.out0 + (static_cast<size_t>(b) * Lres + ri) * P.out0_pitch + P.out0_col0 + h * kAttnD;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(o0 + 8 * c) =
                make_uint4(pack_bf16x2(__uint_as_float(r0[8 * c]) * sc0, __uint_as_float(r0[8 * c + 1]) * sc0),
                           pack_bf16x2(__uint_as_float(r0[8 * c + 2]) * sc0, __uint_as_float(r0[8 * c + 3]) * sc0),
                           pack_bf16x2(__uint_as_float(r0[8 * c + 4]) * sc0, __uint_as_float(r0[8 * c + 5]) * sc0),
                           pack_bf16x2(__uint_as_float(r0[8 * c + 6]) * sc0, __uint_as_float(r0[8 * c + 7]) * sc0));
          if (kDkv) {
            __nv_bfloat16* o1 = P.out1 + (static_cast<size_t>(b) * Lres + ri) * P.out1_pitch + P.out1_col0 + h * kAttnD;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(o1 + 8 * c) =
                  make_uint4(pack_bf16x2(__uint_as_float(r1[8 * c]), __uint_as_float(r1[8 * c + 1])),
                             pack_bf16x2(__uint_as_float(r1[8 * c + 2]), __uint_as_float(r1[8 * c + 3])),
                             pack_bf16x2(__uint_as_float(r1[8 * c + 4]), __uint_as_float(r1[8 * c + 5])),
                             pack_bf16x2(__uint_as_float(r1[8 * c + 6]), __uint_as_float(r1[8 * c + 7])));
          }
        }
      } else if (row_ok) {
        __nv_bfloat16* o0 = P.out0 + (static_cast<size_t>(b) * Lres + ri) * P.out0_pitch + P.out0_col0 + h * kAttnD;
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(o0 + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
        if (kDkv) {
          __nv_bfloat16* o1 = P.out1 + (static_cast<size_t>(b) * Lres + ri) * P.out1_pitch + P.out1_col0 + h * kAttnD;
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(o1 + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    if (pingpong && t == 0) named_bar_sync(1, 256);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace fnd
