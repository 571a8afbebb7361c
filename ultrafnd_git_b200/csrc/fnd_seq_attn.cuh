// fnd_seq_attn.cuh — multi-head cross-attention forward for the sequence front-end (Tier B), flash-style on tcgen05.
//
//   O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + key_padding_mask[b,k] ) V[b,k,h,:]       d_k = 64
//
// One CTA owns one (sample, head, 128-query tile) at a time and walks the key/value sequence in blocks of 128:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (SWIZZLE_128B, 3-D maps: rows past the sequence end of
//               THIS sample are zero-filled) through a 3-stage ring
//   warp 1      one elected thread issues tcgen05.mma:  S_j = Q K_j^T (128x128x64, into one of two TMEM S buffers — S_{j+1}
//               is issued BEFORE P_j V_j so the tensor pipe works while the softmax warps are busy) and
//               T_j = P_j V_j (128x64x128, P from shared memory, V MN-major, into one of two TMEM buffers)
//   warps 2..9  softmax: two threads per query row (each owns half of the block's 128 key columns and half of the 64 output
//               columns; the row maximum is exchanged through shared memory) — tcgen05.ld of the S row, key-padding mask,
//               running max / sum (online softmax, exp2 with the scale folded in, packed f32x2 arithmetic), P_j as bf16
//               into swizzled shared memory for the next MMA, O accumulated in REGISTERS (acc = acc * alpha_j + T_{j-1}:
//               no TMEM read-modify-write correction pass)
// 128-key blocks halve the issue groups and barrier rounds per key (the block cadence was set by the MMA-issuing warp and
// by fixed synchronisation, profiles/r02_attn_phase_stamps.txt). TMEM: 2 x 128 (S) + 2 x 64 (T) columns, ~181 KB of
// shared memory: one persistent CTA per SM. A query row with no valid key yields zeros (LSE = -inf).
//
// No counterpart in the reference (the reference's "co-attention" is a per-sample sigmoid gate,
// src/models/fusion/cross_modal_transformer.py:39-55); checked against the self-oracle oracle/seq_oracle.py.
#pragma once
#include "fnd_common.cuh"

namespace fnd {

constexpr int kAttnBQ = 128;                 // query rows per CTA
constexpr int kAttnBK = 128;                 // keys per block (two 64-key panels)
constexpr int kAttnD = 64;                   // head dimension
constexpr int kAttnStages = 3;
constexpr int kAttnThreads = 320;             // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..9 softmax (two per TMEM lane quarter)
constexpr int kAttnTmemCols = 512;           // S: 2 x 128 columns, T: 2 x 64 columns (power-of-two allocation)
constexpr int kAttnQBytes = kAttnBQ * kAttnD * 2;        // 16 KB
constexpr int kAttnPBytes = kAttnBQ * kAttnBK * 2;       // 32 KB: two K-major panels of 128 rows x 64 keys
constexpr int kAttnKBytes = kAttnBK * kAttnD * 2;        // 16 KB
constexpr int kAttnSmemBytes = 1024 /*align*/ + 4096 /*barriers + row-statistics exchange*/ + 2 * kAttnQBytes + 2 * kAttnPBytes + kAttnStages * 2 * kAttnKBytes;

struct alignas(64) AttnParams {
  CUtensorMap tmQ, tmK, tmV;                 // 3-D [batch][rows][cols] maps (fnd_tmap.h: encode_bf16_3d)
  int B, H, Lq, Lk;
  int q_col0, k_col0, v_col0;                // first column of head 0 inside the Q / K / V matrices
  const int* kv_len;                         // [B] valid prefix length of the key sequence, or null (= Lk)
  const unsigned char* kv_mask;              // [B, Lk] 1 = valid key, or null; combined with kv_len
  float scale_log2;                          // (1/sqrt(d_k)) * log2(e)
  float scale;
  __nv_bfloat16* out;                        // [B*Lq, out_pitch], head h at columns h*64
  int out_pitch;
  float* lse;                                // [B, H, Lq] natural-log logsumexp of the scaled scores, or null
  int* err;
  long long* dbg;                            // probe builds only: [grid][8] accumulated phase cycles
};

// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 issue two fp32 operations per instruction) and the 3-input maximum
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kDbg: probe build — lane 0 of softmax warp 2 accumulates clock64() deltas of the phases of its block loop into
// P.dbg[blockIdx.x * 8 + phase] (tools/seq_probe.py --stamps); the production instantiation carries none of it.
#define ATTN_STAMP(i)                                        \
  do {                                                       \
    if (kDbg && dbg_on) {                                    \
      const long long _t = clock64();                        \
      dbg_acc[i] += _t - dbg_t;                              \
      dbg_t = _t;                                            \
    }                                                        \
  } while (0)
template <bool kDbg>
__global__ void __launch_bounds__(kAttnThreads, 1) seq_attn_fwd_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);     // [2]: the Q tile is double-buffered (next item's tile is
  uint64_t* q_empty = q_full + 2;                            //      prefetched a whole item ahead)
  uint64_t* kv_full = q_empty + 2;
  uint64_t* kv_empty = kv_full + kAttnStages;
  uint64_t* s_full = kv_empty + kAttnStages;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* xchg = reinterpret_cast<float*>(smem + 1024);       // [2 blocks][2 halves][128 rows] row maxima + [2][128] row sums
  uint8_t* sQ = smem + 4096;
  uint8_t* sP = sQ + 2 * kAttnQBytes;
  uint8_t* sKV = sP + 2 * kAttnPBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PERSISTENT: a CTA walks work items w = (sample, head, 128-query tile); consecutive w share the sample and head, so the
  // CTAs running at the same time read the same K / V tiles from L2. All barrier phases and buffer indices run on GLOBAL
  // counters across items, so the producer's K / V ring and the S stream run ahead into the next item (the CTA set-up and
  // the first-tile load latency — ~20 % of a one-item CTA's life, ncu — are paid once per CTA instead of once per item).
  const int nqt = (P.Lq + kAttnBQ - 1) / kAttnBQ;
  const int nwork = nqt * P.H * P.B;
  auto item_nblk = [&](int w) -> int {
    const int b = w / (nqt * P.H);
    int kv_len = P.Lk;
    if (P.kv_len) kv_len = min(max(P.kv_len[b], 0), P.Lk);
    return (kv_len + kAttnBK - 1) / kAttnBK;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
      for (int s = 0; s < kAttnStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 8);                // one arrival per softmax warp (lane 0 after __syncwarp): 256 per-thread
        mbar_init(&p_full[i], 8);                // arrivals per barrier and block flooded the MIO queue (ncu: mio_throttle)
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 8);
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

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1u;                          // parity to wait for on kv_empty (first pass: fresh barrier, passes)
      uint32_t qn = 0;                           // non-empty items so far
      bool ok = true;
#pragma unroll 1
      for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
        const int nblk = item_nblk(w);
        if (nblk == 0) continue;
        const int qt = w % nqt, h = (w / nqt) % P.H, b = w / (nqt * P.H);
        // the last S = Q K^T of the item that used this Q buffer (two items ago) has retired before it is overwritten
        ok = mbar_wait_fast(&q_empty[qn & 1u], ((qn >> 1) & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_PRODUCER);
        if (!ok) break;
        mbar_arrive_expect_tx(&q_full[qn & 1u], kAttnQBytes);
        tma_load_3d(sQ + (qn & 1u) * kAttnQBytes, &P.tmQ, &q_full[qn & 1u], P.q_col0 + h * kAttnD, qt * kAttnBQ, b, kEvictFirst);
        ++qn;
        const int kc = P.k_col0 + h * kAttnD, vc = P.v_col0 + h * kAttnD;
#pragma unroll 1
        for (int j = 0; j < nblk; ++j) {
          ok = mbar_wait_fast(&kv_empty[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
          if (!ok) break;
          mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnKBytes);
          uint8_t* sK = sKV + s * 2 * kAttnKBytes;
          tma_load_3d(sK, &P.tmK, &kv_full[s], kc, j * kAttnBK, b, kEvictLast);
          tma_load_3d(sK + kAttnKBytes, &P.tmV, &kv_full[s], vc, j * kAttnBK, b, kEvictLast);
          if (++s == kAttnStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Warp-uniform control flow: every lane follows the barriers, ONE elected lane issues (under a plain `lane == 0`
    // branch the compiler wraps every tcgen05.mma operand in a divergence-safe uniform-register loop, ~25 instructions each).
    const uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBK, 0, 0);      // S = Q K^T : both K-major
    const uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnD, 0, 1);       // T = P V   : V is [keys][d] = MN-major B
    const uint32_t dhi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16);
    const uint32_t p_lo = smem_desc_lo(smem_u32(sP), 16);                  // + (g & 1) * (kAttnPBytes >> 4)
    const uint32_t k_lo = smem_desc_lo(smem_u32(sKV), 16);                 // + stage * (2 * kAttnKBytes >> 4)
    const uint32_t v_lo = smem_desc_lo(smem_u32(sKV + kAttnKBytes), 8192);
    bool ok = true;
    uint32_t gs = 0, gp = 0;                                               // blocks whose S / P V have been issued (global)
    int ss = 0; uint32_t sph = 0u;                                         // kv stage / parity of the S stream
    int ps = 0;                                                            // kv stage of the P V stream
    uint32_t qn = 0, qcur = 0;                                             // items started by the S stream; Q buffer in use
    auto issue_s = [&](bool last_of_item) {
      ok = ok && mbar_wait_fast(&kv_full[ss], sph, P.err, FND_DEV_TIMEOUT_MMA);
      ok = ok && mbar_wait_fast(&s_free[gs & 1u], ((gs >> 1) & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
      tc_fence_after_sync();
      const uint32_t kl = k_lo + static_cast<uint32_t>(ss) * ((2 * kAttnKBytes) >> 4);
      const uint32_t tS = tmem_base + (gs & 1u) * kAttnBK;
      if (ok && elect_one()) {
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_f16(tS, desc64(q_lo + qcur * (kAttnQBytes >> 4) + 2 * k, dhi), desc64(kl + 2 * k, dhi), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[gs & 1u]);
        if (last_of_item) umma_commit(&q_empty[qcur]);                     // this Q buffer may be refilled once this retires
      }
      __syncwarp();
      ++gs;
      if (++ss == kAttnStages) { ss = 0; sph ^= 1u; }
    };
    // The S stream runs ONE BLOCK AHEAD of the P V stream, across item boundaries too: while the softmax warps work on the
    // last block of an item, S of the next item's first block is already issued (its Q tile was requested when the last S
    // of this item retired), so an item change costs no tensor-pipe bubble.
    int ws = blockIdx.x, js = 0, ns = 0;                                   // S stream: item, block, blocks of that item
    auto s_next_item = [&]() {                                             // advance to the next non-empty item (ns = 0: none left)
      for (; ws < nwork; ws += gridDim.x) {
        ns = item_nblk(ws);
        if (ns > 0) return;
      }
      ns = 0;
    };
    auto s_step = [&]() {                                                  // issue S for (ws, js), advance the S stream
      if (ns == 0) return;
      if (js == 0) {
        qcur = qn & 1u;
        ok = ok && mbar_wait_fast(&q_full[qcur], (qn >> 1) & 1u, P.err, FND_DEV_TIMEOUT_MMA);
        ++qn;
      }
      issue_s(js + 1 == ns);
      if (++js == ns) { js = 0; ws += gridDim.x; s_next_item(); }
    };
    s_next_item();
    s_step();
#pragma unroll 1
    for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
      const int nblk = item_nblk(w);
#pragma unroll 1
      for (int j = 0; j < nblk && ok; ++j) {
        s_step();                                                          // S of the block after this one (same or next item)
        const uint32_t par = (gp >> 1) & 1u;
        ok = ok && mbar_wait_fast(&p_full[gp & 1u], par, P.err, FND_DEV_TIMEOUT_MMA);
        ok = ok && mbar_wait_fast(&o_free[gp & 1u], par ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        tc_fence_after_sync();
        const uint32_t pl = p_lo + (gp & 1u) * (kAttnPBytes >> 4);
        const uint32_t vl = v_lo + static_cast<uint32_t>(ps) * ((2 * kAttnKBytes) >> 4);
        const uint32_t tO = tmem_base + 2 * kAttnBK + (gp & 1u) * kAttnD;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kAttnBK / 16; ++k)       // keys 0..63 from P panel 0, 64..127 from panel 1 (16 KB apart)
            umma_f16(tO, desc64(pl + (k >> 2) * (kAttnPBytes >> 5) + 2 * (k & 3), dhi), desc64(vl + 128 * k, dhi), idesc_o, k != 0 ? 1u : 0u);
          umma_commit(&o_full[gp & 1u]);
          umma_commit(&kv_empty[ps]);
        }
        __syncwarp();
        ++gp;
        if (++ps == kAttnStages) ps = 0;
      }
    }
  } else {
    // ================= softmax + output: warps 2..9 =================
    // Two warps share each TMEM lane quarter (hardware rule: a warp reads lanes 32 * (warp % 4) ...): for its query row a
    // thread owns HALF of the block's 64 key columns and half of the 64 output columns. The row maximum is the only value
    // the two halves exchange per block (shared memory + a 64-thread named barrier); the row sums stay partial until the end.
    const int half = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint64_t sl2 = pack_f32x2(P.scale_log2, P.scale_log2);
    const uint32_t xchg_s = smem_u32(xchg), sP_s = smem_u32(sP);
    bool ok = true;
    uint32_t g = 0;                              // key blocks processed so far, over all items (buffer index / parity)
    uint64_t acc[kAttnD / 4];                    // 32 output columns as packed f32x2
    const bool dbg_on = kDbg && P.dbg != nullptr && warp == 2 && lane == 0;
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long dbg_t = kDbg ? clock64() : 0;

    auto add_tmp = [&](uint32_t gg) {   // acc += T_gg  (this thread's half of the P V product of block gg), then release the buffer
      ok = ok && mbar_wait_fast(&o_full[gg & 1u], (gg >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + lane_addr + 2 * kAttnBK + (gg & 1u) * kAttnD + half * 32, r);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[gg & 1u]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = add_f32x2(acc[i], pack_u32x2(r[2 * i], r[2 * i + 1]));
    };

    // (sample, head, query tile) of the current item, advanced incrementally (no integer divisions on the item path)
    int qt = blockIdx.x % nqt, h = (blockIdx.x / nqt) % P.H, b = blockIdx.x / (nqt * P.H);
#pragma unroll 1
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      int kv_len = P.Lk;
      if (P.kv_len) kv_len = min(max(__ldg(P.kv_len + b), 0), P.Lk);
      const int nblk = (kv_len + kAttnBK - 1) / kAttnBK;
      const int qi = qt * kAttnBQ + row;
      const unsigned char* mrow = P.kv_mask ? P.kv_mask + static_cast<size_t>(b) * P.Lk : nullptr;
      // off_run = running row maximum in exp2 units (m * scale * log2 e), -inf while no valid key has been seen
      float off_run = -INFINITY, l_part = 0.f;
#pragma unroll
      for (int i = 0; i < kAttnD / 4; ++i) acc[i] = 0ull;

#pragma unroll 1
      for (int j = 0; j < nblk; ++j, ++g) {
        ATTN_STAMP(7);                             // loop overhead / item epilogue
        ok = ok && mbar_wait_fast(&s_full[g & 1u], (g >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        ATTN_STAMP(0);                             // wait for S
        tc_fence_after_sync();
        float s[64];
        {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + (g & 1u) * kAttnBK + half * 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
          tmem_ld_32x32(tmem_base + lane_addr + (g & 1u) * kAttnBK + half * 64 + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[32 + i] = __uint_as_float(r[i]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g & 1u]);
        ATTN_STAMP(1);                             // TMEM load of S + release

        // ---- key-padding mask of this thread's 64 columns as two validity words (warp-cooperative: two ballots) ----
        const int k0 = j * kAttnBK + half * 64;
        if (mrow || k0 + 64 > kv_len) {
          const int ka = k0 + lane, kb = ka + 32;
          const uint32_t va = __ballot_sync(0xffffffffu, ka < kv_len && (!mrow || mrow[ka] != 0));
          const uint32_t vb = __ballot_sync(0xffffffffu, kb < kv_len && (!mrow || mrow[kb] != 0));
          if ((va & vb) != 0xffffffffu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = ((va >> i) & 1u) ? s[i] : -INFINITY;
              s[32 + i] = ((vb >> i) & 1u) ? s[32 + i] : -INFINITY;
            }
          }
        }
        // ---- row maximum: 3-input max tree over 64 columns (four independent chains), then the other half's maximum ----
        float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]), m2 = fmax3(s[6], s[7], s[8]), m3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
        for (int i = 12; i < 60; i += 8) {
          m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
          m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
        }
        m0 = fmax3(m0, s[60], s[61]); m1 = fmax3(m1, s[62], s[63]);
        float mx = fmaxf(fmax3(m0, m1, m2), m3);
        const uint32_t xc = xchg_s + (g & 1u) * 1024u + static_cast<uint32_t>(row * 4);
        sts_f32(xc + half * 512, mx);
        pair_bar_sync(1 + qd);
        mx = fmaxf(mx, lds_f32(xc + (half ^ 1) * 512));
        ATTN_STAMP(2);                             // mask + row maximum + exchange

        const float off_new = fmaxf(off_run, mx * P.scale_log2);              // scale_log2 > 0: max commutes with the scaling
        const float off = (off_new == -INFINITY) ? 0.f : off_new;
        const float alpha = ex2_approx(off_run - off);                        // off_run = -inf -> 0; unchanged maximum -> exactly 1
        off_run = off_new;
        const uint64_t noff = pack_f32x2(-off, -off);
        uint32_t pk[32];
        uint64_t ps0 = 0ull, ps1 = 0ull;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          float x0, x1, x2, x3;
          unpack_f32x2(fma_f32x2(pack_f32x2(s[i], s[i + 1]), sl2, noff), x0, x1);
          unpack_f32x2(fma_f32x2(pack_f32x2(s[i + 2], s[i + 3]), sl2, noff), x2, x3);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
          ps0 = add_f32x2(ps0, pack_f32x2(p0, p1));
          ps1 = add_f32x2(ps1, pack_f32x2(p2, p3));
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        {
          float a0, a1;
          unpack_f32x2(add_f32x2(ps0, ps1), a0, a1);
          l_part = fmaf(l_part, alpha, a0 + a1);
        }
        ATTN_STAMP(3);                             // exp2 + row sum + bf16 pack
        // ---- P -> shared memory in the SWIZZLE_128B K-major layout the MMA descriptor expects: this thread's 64 keys are
        //      one full 128-byte row of panel `half`: row r at r*128 B, 16-byte chunk c stored at chunk (c ^ (r & 7)) ----
        {
          const uint32_t prow = sP_s + (g & 1u) * kAttnPBytes + static_cast<uint32_t>(half * (kAttnPBytes >> 1) + row * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts_v4(prow + static_cast<uint32_t>((c ^ (row & 7)) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1u]);
        ATTN_STAMP(4);                             // P store + proxy fence + arrive
        // ---- fold in the previous block's product while the tensor pipe works on this one ----
        if (j > 0) add_tmp(g - 1);
        ATTN_STAMP(5);                             // wait for / load / add the previous P V product
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          const uint64_t a2 = pack_f32x2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = mul_f32x2(acc[i], a2);
        }
        ATTN_STAMP(6);                             // rescale
      }
      float l_run = 0.f;
      if (nblk > 0) {
        add_tmp(g - 1);
        // total row sum = the two halves' partial sums
        const uint32_t xl = xchg_s + 2048u + static_cast<uint32_t>(row * 4);
        sts_f32(xl + half * 512, l_part);
        pair_bar_sync(1 + qd);
        l_run = l_part + lds_f32(xl + (half ^ 1) * 512);
      }
      if (qi < P.Lq) {
        const float inv = (ok && l_run > 0.f) ? __fdividef(1.f, l_run) : 0.f;
        __nv_bfloat16* op = P.out + (static_cast<size_t>(b) * P.Lq + qi) * P.out_pitch + h * kAttnD + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float o[8];
#pragma unroll
          for (int t = 0; t < 4; ++t) unpack_f32x2(acc[4 * c + t], o[2 * t], o[2 * t + 1]);
          *reinterpret_cast<uint4*>(op + 8 * c) = make_uint4(pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                                                             pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
        }
        if (P.lse && half == 0)
          P.lse[(static_cast<size_t>(b) * P.H + h) * P.Lq + qi] = (l_run > 0.f) ? fmaf(off_run, 0.69314718055994531f, __logf(l_run)) : -INFINITY;
      }
      qt += static_cast<int>(gridDim.x);
      while (qt >= nqt) {
        qt -= nqt;
        if (++h == P.H) { h = 0; ++b; }
      }
    }
    if (kDbg && dbg_on) {
#pragma unroll
      for (int i = 0; i < 8; ++i) P.dbg[static_cast<size_t>(blockIdx.x) * 8 + i] = dbg_acc[i];
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}
#undef ATTN_STAMP

}  // namespace fnd
