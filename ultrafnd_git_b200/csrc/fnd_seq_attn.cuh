// fnd_seq_attn.cuh — multi-head cross-attention forward for the sequence front-end (Tier B), flash-style on tcgen05.
//
//   O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + key_padding_mask[b,k] ) V[b,k,h,:]       d_k = 64
//
// One persistent CTA per SM walks work items (sample, head, 256-query tile). An item is TWO 128-row query tiles that
// share every K / V block; each tile has its own softmax warpgroup, its own S / P / O regions in TMEM, and the two
// warpgroups run half a period apart, so while one sits in the exp2 phase (MUFU-bound: 128 x 128 exponentials per
// block against 512 cycles of MMA) the other does its TMEM loads, row maximum, packing and stores:
//   warps 0..3   softmax, tile 0: ONE thread per query row (TMEM lane = row: no cross-thread reduction at all) —
//                tcgen05.ld of the 128-column S row, key-padding mask, running max / sum (online softmax, exp2 with the
//                scale folded in, packed f32x2 arithmetic), P as bf16 STRAIGHT BACK INTO TMEM (tcgen05.st): the P V
//                product takes its A operand from TMEM, so P never touches shared memory
//   warps 4..7   the same for tile 1
//   warp  8      TMA producer: the item's two Q tiles (double-buffered across items), then K_j / V_j (SWIZZLE_128B,
//                3-D maps: rows past the sequence end of THIS sample are zero-filled) through a 3-stage ring
//   warps 9, 10  MMA issuers, ONE PER TILE (an elected thread each): S_t = Q_t K_j^T (128x128x64) and O_t (+)= P_t V_j
//                (128x64x128, A from TMEM, V MN-major). Each warp follows its own tile's events with blocking waits —
//                s_free_t(j) -> S_t(j+1),  p_full_t(j) -> P_t V(j) — so the tiles never wait for each other's turn in an
//                issuing warp; the shared K / V / Q buffers are released by one tcgen05.commit arrival from each warp
//   warp  11     idle (completes the third warpgroup, which hands its registers to the softmax warpgroups)
// O accumulates in TMEM (the MMA's own accumulate flag). The running maximum is updated LAZILY: only when a row's new
// maximum exceeds the one in use by more than 8 (log2 units) does the warp rescale its O rows in TMEM (tcgen05.ld / mul /
// tcgen05.st); otherwise the stale maximum stays — the probabilities are then at most 2^8, harmless in bf16 / fp32, and
// the final normalisation and LSE are exact either way.
// S_t(j+1) is issued as soon as the softmax threads HOLD S_t(j) in registers (s_free), a whole softmax block before they
// need it, so Q K^T (issue + execution + two barrier hops: ~1100 cycles by the event stamps, tools/attn_stamps.py) is off
// their critical path; pv_done_t orders "P_t V(j) retired" before P_t / O_t are written for block j+1. Every wait except
// the critical one is done ahead of time (an mbarrier try_wait costs ~90 cycles even on a completed phase). The two
// warpgroups take turns in the MUFU-bound exp2 section (named-barrier token). TMEM: S 2 x 128, O 2 x 64, P 2 x 64 (bf16
// pairs) = 512 columns; shared memory 2 x 32 KB Q + 3 x (16 + 16) KB K / V + 2 x 16 KB output staging. A query row with
// no valid key yields zeros (LSE = -inf).
//
// No counterpart in the reference (the reference's "co-attention" is a per-sample sigmoid gate,
// src/models/fusion/cross_modal_transformer.py:39-55); checked against the self-oracle oracle/seq_oracle.py.
#pragma once
#include "fnd_common.cuh"

namespace fnd {

constexpr int kAttnBQ = 128;                 // query rows per tile
constexpr int kAttnItemQ = 256;              // query rows per work item (two tiles)
constexpr int kAttnBK = 128;                 // keys per block
constexpr int kAttnD = 64;                   // head dimension
constexpr int kAttnStages = 3;
constexpr int kAttnThreads = 384;            // warps 0..7 softmax (two warpgroups), 8 TMA, 9 / 10 MMA (one per tile; 9 owns TMEM), 11 idle
constexpr int kAttnTmemCols = 512;
constexpr int kAttnQBytes = kAttnBQ * kAttnD * 2;        // 16 KB per tile
constexpr int kAttnKBytes = kAttnBK * kAttnD * 2;        // 16 KB
constexpr int kAttnOutBytes = kAttnBQ * kAttnD * 2;      // 16 KB: one 128 x 64 bf16 output tile staged for its TMA store
constexpr int kAttnSmemBytes = 1024 /*align*/ + 1024 /*barriers*/ + 2 * 2 * kAttnQBytes + kAttnStages * 2 * kAttnKBytes + 2 * kAttnOutBytes;
constexpr int kAttnTmemS = 0, kAttnTmemO = 256, kAttnTmemP = 384;
constexpr float kAttnLazyLog2 = 8.f;         // rescale only when the row maximum grew by more than this (log2 units)
constexpr int kAttnDefaultPoly = 0;           // FND_ATTN_POLY overrides (0..3 of every 4 pairs on the FMA pipe)
constexpr int kAttnDefaultSkewNs = 0;         // FND_ATTN_SKEW_NS overrides
constexpr int kAttnDefaultPingPong = 1;       // FND_ATTN_PINGPONG overrides

struct alignas(64) AttnParams {
  CUtensorMap tmQ, tmK, tmV;                 // 3-D [batch][rows][cols] maps (fnd_tmap.h: encode_bf16_3d), box 64 x 128 x 1
  CUtensorMap tmO;                           // the output the same way: rows >= Lq of a tile are clipped by the store
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
  int skew_ns;                               // start-up delay of softmax warpgroup 1 (experiment knob)
  int pingpong;                              // 1: the two softmax warpgroups take turns in the exp2 phase (named barriers)
  int dbg_noexp;                             // probe: skip the exponentials (P = 0) to expose the pure pipeline latency
  long long* dbg;                            // probe: [grid][32 steps][16 events] clock64 stamps (fnd_seq_debug_attn_stamps)
};
#define ATTN_EV(step, ev) do { if (P.dbg && (step) < 32u) P.dbg[(static_cast<size_t>(blockIdx.x) * 32 + (step)) * 16 + (ev)] = clock64(); } while (0)

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
// named barriers (ids 1, 2; id 0 is __syncthreads) shared by the two softmax warpgroups: 128 syncing + 128 arriving threads
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a PAIR on the FMA pipe (no MUFU): n = round(x) by the 1.5 * 2^23 trick, f = x - n in [-0.5, 0.5], 2^f by a cubic
// (max rel. error 7.5e-5, invisible after the bf16 rounding of P), 2^n by adding n to the exponent field. The exp2 phase
// of the softmax is MUFU-bound (16 ex2 / clk / SM against 128 x 128 exponentials per block): part of every row goes this way.
// x is clamped at -120 (a masked / far-away score becomes 2^-120 instead of wrapping the exponent field).
__device__ __forceinline__ void exp2_poly_x2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
  x2 = pack_f32x2(x0, x1);
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f), nmagic = pack_f32x2(-12582912.f, -12582912.f);
  const uint64_t t2 = add_f32x2(x2, magic);
  const uint64_t n2 = add_f32x2(t2, nmagic);
  const uint64_t f2 = fma_f32x2(n2, pack_f32x2(-1.f, -1.f), x2);
  uint64_t q2 = fma_f32x2(pack_f32x2(0.0551716573536396f, 0.0551716573536396f), f2, pack_f32x2(0.2426111400127411f, 0.2426111400127411f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
  uint32_t t0, t1, q0, q1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t2));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(q0), "=r"(q1) : "l"(q2));
  p0 = __uint_as_float(q0 + (t0 << 23));
  p1 = __uint_as_float(q1 + (t1 << 23));
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 rows = lanes, bf16 pairs along the columns, 8 columns per
// K = 16 step) is read from tensor memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 / 32 consecutive 32-bit columns, registers -> tensor memory
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// register re-partitioning between warpgroups (all four warps of a warpgroup execute the same instruction)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int N> struct IntTag { static constexpr int value = N; };
// kPoly: of every 8 exponentials of an UNMASKED block, 2 * kPoly are evaluated on the FMA pipe (exp2_poly_x2), the rest on
// the MUFU. Blocks that carry a mask use the MUFU throughout (a masked score must give exactly 0).
template <int kPoly>
__global__ void __launch_bounds__(kAttnThreads, 1) seq_attn_fwd_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);     // [2]: an item's two Q tiles, double-buffered across items
  uint64_t* q_empty = q_full + 2;
  uint64_t* k_full = q_empty + 2;
  uint64_t* k_empty = k_full + kAttnStages;
  uint64_t* v_full = k_empty + kAttnStages;
  uint64_t* v_empty = v_full + kAttnStages;
  uint64_t* s_full = v_empty + kAttnStages;                  // [2]: per tile
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint64_t* s_free = o_free + 2;                             // [2]: the softmax threads hold S_t in registers
  uint64_t* pv_done = s_free + 2;                            // [2]: P_t V of the previous block has retired (P_t / O_t may be written)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  uint8_t* sQ = smem + 1024;                                 // [2 buffers][2 tiles][128 x 64]
  uint8_t* sKV = sQ + 2 * 2 * kAttnQBytes;                   // [stages][K | V]
  uint8_t* sOut = sKV + kAttnStages * 2 * kAttnKBytes;       // [2 tiles][128 x 64 bf16, SWIZZLE_128B]: output staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items w = (sample, head, 256-query tile), query tile fastest: the CTAs running at the same time read the same
  // K / V tiles from L2. All barrier phases and ring indices run on GLOBAL counters across items, so the producer's ring
  // and the S stream run ahead into the next item.
  const int nqt = (P.Lq + kAttnItemQ - 1) / kAttnItemQ;
  const int nwork = nqt * P.H * P.B;
  auto item_nblk = [&](int w) -> int {
    const int b = w / (nqt * P.H);
    int kv_len = P.Lk;
    if (P.kv_len) kv_len = min(max(P.kv_len[b], 0), P.Lk);
    return (kv_len + kAttnBK - 1) / kAttnBK;
  };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
    tma_prefetch_desc(&P.tmO);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2); }      // "empty": one commit per MMA warp
      for (int s = 0; s < kAttnStages; ++s) {
        mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 2);
        mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 2);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);                // one arrival per softmax warp of the tile (lane 0 after __syncwarp)
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 4);
        mbar_init(&s_free[i], 4);
        mbar_init(&pv_done[i], 1);
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
        uint32_t ph = 1u;                          // parity to wait for on k_empty / v_empty (first pass: fresh barrier, passes)
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
          mbar_arrive_expect_tx(&q_full[qn & 1u], 2 * kAttnQBytes);
          uint8_t* q = sQ + (qn & 1u) * 2 * kAttnQBytes;
          tma_load_3d(q, &P.tmQ, &q_full[qn & 1u], P.q_col0 + h * kAttnD, qt * kAttnItemQ, b, kEvictFirst);
          tma_load_3d(q + kAttnQBytes, &P.tmQ, &q_full[qn & 1u], P.q_col0 + h * kAttnD, qt * kAttnItemQ + kAttnBQ, b, kEvictFirst);
          ++qn;
          const int kc = P.k_col0 + h * kAttnD, vc = P.v_col0 + h * kAttnD;
#pragma unroll 1
          for (int j = 0; j < nblk; ++j) {
            uint8_t* sK = sKV + s * 2 * kAttnKBytes;
            ok = mbar_wait_fast(&k_empty[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
            if (!ok) break;
            mbar_arrive_expect_tx(&k_full[s], kAttnKBytes);
            tma_load_3d(sK, &P.tmK, &k_full[s], kc, j * kAttnBK, b, kEvictLast);
            ok = mbar_wait_fast(&v_empty[s], ph, P.err, FND_DEV_TIMEOUT_PRODUCER);
            if (!ok) break;
            mbar_arrive_expect_tx(&v_full[s], kAttnKBytes);
            tma_load_3d(sK + kAttnKBytes, &P.tmV, &v_full[s], vc, j * kAttnBK, b, kEvictLast);
            if (++s == kAttnStages) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ================= MMA issuers: one warp PER TILE (warp 9: tile 0, warp 10: tile 1) =================
      // Each warp runs its tile's own event sequence  s_free_t(g) -> S_t(g+1),  p_full_t(g) -> O_t += P_t V(g)  with blocking
      // waits, so the two tiles never wait for each other inside an issuing warp (with ONE warp and a fixed issue order the
      // event stamps showed a published P waiting ~2000 cycles for the other tile's turn). The K / V / Q buffers are shared:
      // their "empty" barriers take one tcgen05.commit arrival from each warp. Warp-uniform control flow, ONE elected lane
      // issues; everything but the critical barrier is waited for ahead of time (an mbarrier try_wait costs ~90 cycles even
      // when the phase completed long ago).
      const int t = warp - 9;
      const uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBK, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnD, 0, 1);       // O += P V  : A from TMEM, V is [keys][d] = MN-major B
      const uint32_t dhi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16) + static_cast<uint32_t>(t) * (kAttnQBytes >> 4);
      const uint32_t k_lo = smem_desc_lo(smem_u32(sKV), 16);                 // + stage * (2 * kAttnKBytes >> 4)
      const uint32_t v_lo = smem_desc_lo(smem_u32(sKV + kAttnKBytes), 8192);
      const uint32_t tS = tmem_base + kAttnTmemS + static_cast<uint32_t>(t) * kAttnBK;
      const uint32_t tO = tmem_base + kAttnTmemO + static_cast<uint32_t>(t) * kAttnD;
      const uint32_t tP = tmem_base + kAttnTmemP + static_cast<uint32_t>(t) * (kAttnBK / 2);
      bool ok = true;
      // ---- S stream cursor: one block AHEAD of the P V stream, across item boundaries ----
      int ws = blockIdx.x, js = 0, ns = 0;                                   // item, block, blocks of that item
      int ks = 0; uint32_t kph = 0u;                                         // K stage / parity
      uint32_t qn = 0, qcur = 0;                                             // items started by the S stream; Q buffer in use
      bool qk_ready = false;
      auto s_next_item = [&]() {                                             // advance to the next non-empty item (ns = 0: none left)
        for (; ws < nwork; ws += gridDim.x) {
          ns = item_nblk(ws);
          if (ns > 0) return;
        }
        ns = 0;
      };
      auto prepare_qk = [&]() {                                              // operand waits of the next S step, done early
        if (ns == 0 || qk_ready) return;
        if (js == 0) {
          qcur = qn & 1u;
          ok = ok && mbar_wait_fast(&q_full[qcur], (qn >> 1) & 1u, P.err, FND_DEV_TIMEOUT_MMA);
          ++qn;
        }
        ok = ok && mbar_wait_fast(&k_full[ks], kph, P.err, FND_DEV_TIMEOUT_MMA);
        qk_ready = true;
      };
      auto issue_qk = [&]() {                                                // S_t of block (ws, js)
        if (ns == 0) return;
        prepare_qk();
        tc_fence_after_sync();
        const uint32_t ql = q_lo + qcur * 2u * (kAttnQBytes >> 4);
        const uint32_t kl = k_lo + static_cast<uint32_t>(ks) * ((2 * kAttnKBytes) >> 4);
        if (ok && elect_one()) {
#pragma unroll
          for (int k = 0; k < kAttnD / 16; ++k)
            umma_f16(tS, desc64(ql + 2 * k, dhi), desc64(kl + 2 * k, dhi), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[ks]);                                         // second arrival comes from the other tile's warp
          if (js + 1 == ns) umma_commit(&q_empty[qcur]);                     // this Q buffer may be refilled once both tiles are through
        }
        __syncwarp();
        if (++ks == kAttnStages) { ks = 0; kph ^= 1u; }
        if (++js == ns) { js = 0; ws += gridDim.x; s_next_item(); }
        qk_ready = false;
      };
      // ---- P V stream cursor ----
      int pw = blockIdx.x, pj = 0, pn = 0;
      uint32_t pg = 0u, pi = 0u;                                             // global step / non-empty item count
      int vs = 0; uint32_t vph = 0u;
      auto pv_next_item = [&]() {
        for (; pw < nwork; pw += gridDim.x) {
          pn = item_nblk(pw);
          if (pn > 0) return;
        }
        pn = 0;
      };
      s_next_item();
      pv_next_item();
      issue_qk();                                                            // S_t of the first block
#pragma unroll 1
      while (ok && pn != 0) {
        // S_t(g+1) as soon as the softmax threads HOLD S_t(g) in registers: a whole softmax block before they need it
        if (ns != 0) {
          prepare_qk();
          ok = ok && mbar_wait_fast(&s_free[t], pg & 1u, P.err, FND_DEV_TIMEOUT_MMA);
          issue_qk();
          if (lane == 0) ATTN_EV(pg, 10 + 4 * t);
        }
        // O_t (+)= P_t V(g): V and (at an item start) the drained O are waited for before the critical wait on P
        const int j = pj;
        ok = ok && mbar_wait_fast(&v_full[vs], vph, P.err, FND_DEV_TIMEOUT_MMA);
        if (j == 0) ok = ok && mbar_wait_fast(&o_free[t], (pi & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        ok = ok && mbar_wait_fast(&p_full[t], pg & 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (lane == 0) ATTN_EV(pg, 8 + 4 * t);
        tc_fence_after_sync();
        const uint32_t vl = v_lo + static_cast<uint32_t>(vs) * ((2 * kAttnKBytes) >> 4);
        if (ok && elect_one()) {
#pragma unroll
          for (int k = 0; k < kAttnBK / 16; ++k)
            umma_f16_ts(tO, tP + 8 * k, desc64(vl + 128 * k, dhi), idesc_o, (j != 0 || k != 0) ? 1u : 0u);
          umma_commit(&pv_done[t]);
          if (j + 1 == pn) umma_commit(&o_full[t]);
          umma_commit(&v_empty[vs]);                                         // second arrival comes from the other tile's warp
        }
        __syncwarp();
        if (lane == 0) ATTN_EV(pg, 9 + 4 * t);
        ++pg;
        if (++vs == kAttnStages) { vs = 0; vph ^= 1u; }
        if (++pj == pn) { pj = 0; ++pi; pw += gridDim.x; pv_next_item(); }
      }
    }
  } else {
    // ================= softmax + output: warps 0..7, one thread per query row =================
    setmaxnreg_inc<216>();
    const int t = warp >> 2;                     // tile of the item
    const int qd = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + kAttnTmemS + static_cast<uint32_t>(t) * kAttnBK;
    const uint32_t tO = tmem_base + lane_addr + kAttnTmemO + static_cast<uint32_t>(t) * kAttnD;
    const uint32_t tP = tmem_base + lane_addr + kAttnTmemP + static_cast<uint32_t>(t) * (kAttnBK / 2);
    const uint64_t sl2 = pack_f32x2(P.scale_log2, P.scale_log2);
    bool ok = true;
    uint32_t g = 0, ip = 0;                      // key blocks / non-empty items processed so far (barrier parities)

    if (t == 1 && P.skew_ns > 0) __nanosleep(static_cast<unsigned>(P.skew_ns));
    // PING-PONG: the exp2 phase is MUFU-bound and everything else a softmax thread does (waiting for S, tcgen05.ld, row
    // maximum, tcgen05.st, the MMA round trip) is not, so the two warpgroups must be in the exp2 phase at DIFFERENT times.
    // Left alone they fall into lockstep (ncu: both in the exp2 section at once, then both waiting for S). A token passed
    // through two named barriers serialises the exp2 phases: group t enters after `bar.sync 1 + t`, leaves with
    // `bar.arrive 2 - t`. Group 1 hands group 0 the first token; group 0 absorbs the last one before the CTA ends.
    const bool pingpong = P.pingpong != 0;
    if (pingpong && t == 1) named_bar_arrive(1, 256);
#pragma unroll 1
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int qt = w % nqt, h = (w / nqt) % P.H, b = w / (nqt * P.H);
      int kv_len = P.Lk;
      if (P.kv_len) kv_len = min(max(__ldg(P.kv_len + b), 0), P.Lk);
      const int nblk = (kv_len + kAttnBK - 1) / kAttnBK;
      const int qi = qt * kAttnItemQ + t * kAttnBQ + row;
      const unsigned char* mrow = P.kv_mask ? P.kv_mask + static_cast<size_t>(b) * P.Lk : nullptr;
      // m_run = row maximum in use, in exp2 units (m * scale * log2 e), -inf while no valid key has been seen
      float m_run = -INFINITY, l_run = 0.f;

#pragma unroll 1
      for (int j = 0; j < nblk; ++j, ++g) {
        ok = ok && mbar_wait_fast(&s_full[t], g & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        if (qd == 0 && lane == 0) ATTN_EV(g, 0 + 4 * t);
        tc_fence_after_sync();
        float s[kAttnBK];
#pragma unroll
        for (int c = 0; c < kAttnBK / 32; ++c) tmem_ld_32x32(tS + 32 * c, reinterpret_cast<uint32_t(&)[32]>(s[32 * c]));
        tmem_ld_wait();
        // the whole S row is in registers: the MMA warp may overwrite S_t with the next block's scores right away
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if (qd == 0 && lane == 0) ATTN_EV(g, 1 + 4 * t);

        // ---- key-padding mask of the block's 128 columns as four validity words (warp-cooperative ballots) ----
        const int k0 = j * kAttnBK;
        const bool masked_blk = mrow || k0 + kAttnBK > kv_len;
        if (masked_blk) {
#pragma unroll
          for (int c = 0; c < kAttnBK / 32; ++c) {
            const int kk = k0 + 32 * c + lane;
            const uint32_t v = __ballot_sync(0xffffffffu, kk < kv_len && (!mrow || mrow[kk] != 0));
            if (v != 0xffffffffu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) s[32 * c + i] = ((v >> i) & 1u) ? s[32 * c + i] : -INFINITY;
            }
          }
        }
        // ---- row maximum: 3-input max tree over 128 columns (four independent chains) ----
        float m0 = fmax3(s[0], s[1], s[2]), m1 = fmax3(s[3], s[4], s[5]), m2 = fmax3(s[6], s[7], s[8]), m3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
        for (int i = 12; i < 124; i += 8) {
          m0 = fmax3(m0, s[i], s[i + 1]); m1 = fmax3(m1, s[i + 2], s[i + 3]);
          m2 = fmax3(m2, s[i + 4], s[i + 5]); m3 = fmax3(m3, s[i + 6], s[i + 7]);
        }
        m0 = fmax3(m0, s[124], s[125]); m1 = fmax3(m1, s[126], s[127]);
        const float mx = fmaxf(fmax3(m0, m1, m2), m3);
        const float m_new = fmaxf(m_run, mx * P.scale_log2);                  // scale_log2 > 0: max commutes with the scaling
        if (j > 0) {
          // P_t V of block j-1 must have retired before P_t is overwritten / O_t rescaled (issued a whole softmax block ago).
          // For j == 0 the item epilogue's o_full wait covered it.
          ok = ok && mbar_wait_fast(&pv_done[t], (g - 1u) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
          tc_fence_after_sync();
        }
        if (j == 0) {
          m_run = m_new;                                                      // O is overwritten by the first P V: nothing to rescale
        } else if (__any_sync(0xffffffffu, m_new - m_run > kAttnLazyLog2)) {
          // lazy rescale (warp-uniform branch): P V of block j-1 has retired (s_full of block j was committed after it),
          // and P V of block j is issued only after this thread's p_full arrival below
          const float alpha = (m_new == -INFINITY) ? 1.f : ex2_approx(m_run - m_new);   // m_run = -inf -> 0 (O is 0 there)
          m_run = m_new;
          l_run *= alpha;
          const uint64_t a2 = pack_f32x2(alpha, alpha);
#pragma unroll
          for (int c = 0; c < kAttnD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tO + 32 * c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t v = mul_f32x2(pack_u32x2(r[2 * i], r[2 * i + 1]), a2);
              asm("mov.b64 {%0, %1}, %2;" : "=r"(r[2 * i]), "=r"(r[2 * i + 1]) : "l"(v));
            }
            tmem_st_32x32(tO + 32 * c, r);
          }
        }
        const float off = (m_run == -INFINITY) ? 0.f : m_run;
        const uint64_t noff = pack_f32x2(-off, -off);
        uint64_t ps0 = 0ull, ps1 = 0ull;
        // ---- p = exp2(s * scale_log2 - off): 32 columns at a time -> 16 bf16 pairs -> tcgen05.st into the P region ----
        auto exp_block = [&](auto poly_tag) {
          constexpr int kP = decltype(poly_tag)::value;        // pairs per 8 columns that go to the FMA pipe
#pragma unroll
          for (int c = 0; c < kAttnBK / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float x0, x1, x2, x3, p0, p1, p2, p3;
              const uint64_t xa = fma_f32x2(pack_f32x2(s[32 * c + i], s[32 * c + i + 1]), sl2, noff);
              const uint64_t xb = fma_f32x2(pack_f32x2(s[32 * c + i + 2], s[32 * c + i + 3]), sl2, noff);
              if ((i & 4) ? (kP >= 3) : (kP >= 1)) exp2_poly_x2(xa, p0, p1);
              else { unpack_f32x2(xa, x0, x1); p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
              if ((i & 4) ? (kP >= 4) : (kP >= 2)) exp2_poly_x2(xb, p2, p3);
              else { unpack_f32x2(xb, x2, x3); p2 = ex2_approx(x2); p3 = ex2_approx(x3); }
              ps0 = add_f32x2(ps0, pack_f32x2(p0, p1));
              ps1 = add_f32x2(ps1, pack_f32x2(p2, p3));
              pk[i >> 1] = pack_bf16x2(p0, p1);
              pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            tmem_st_32x16(tP + 16 * c, pk);
          }
        };
        if (pingpong) named_bar_sync(1 + t, 256);
        if (qd == 0 && lane == 0) ATTN_EV(g, 2 + 4 * t);
        if (P.dbg_noexp) {
          uint32_t z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(s[i]) & 0u;
#pragma unroll
          for (int c = 0; c < kAttnBK / 32; ++c) tmem_st_32x16(tP + 16 * c, z);
        } else if (kPoly == 0 || masked_blk) exp_block(IntTag<0>{});   // a masked score must give exactly 0: MUFU throughout
        else exp_block(IntTag<kPoly>{});
        if (pingpong) named_bar_arrive(2 - t, 256);
        {
          float a0, a1;
          unpack_f32x2(add_f32x2(ps0, ps1), a0, a1);
          l_run += a0 + a1;
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (qd == 0 && lane == 0) ATTN_EV(g, 3 + 4 * t);
      }

      // ---- item epilogue: O / l -> bf16 -> swizzled shared-memory tile -> ONE TMA store per tile. (Row stores straight
      //      from registers made every warp-level store touch 32 different 128-byte lines: 256 LSU wavefronts per warp and
      //      item, ~11 % of the softmax warps' time in ncu.) Rows past Lq are clipped by the tensor map. ----
      uint32_t r[kAttnD];
      float inv = 0.f;
      if (nblk > 0) {
        ok = ok && mbar_wait_fast(&o_full[t], ip & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
        tc_fence_after_sync();
        tmem_ld_32x32(tO, reinterpret_cast<uint32_t(&)[32]>(r[0]));
        tmem_ld_32x32(tO + 32, reinterpret_cast<uint32_t(&)[32]>(r[32]));
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
        ++ip;
        inv = (ok && l_run > 0.f) ? __fdividef(1.f, l_run) : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < kAttnD; ++i) r[i] = 0u;       // no valid key at all: zeros
      }
      uint8_t* stg = sOut + t * kAttnOutBytes;
      if (qd == 0 && lane == 0) tma_store_wait_read<0>();    // this tile's previous store has drained the staging buffer
      named_bar_sync(3 + t, 128);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(r[8 * c + i]) * inv;
        *reinterpret_cast<uint4*>(stg + row * 128 + ((c ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
      fence_proxy_async_smem();
      named_bar_sync(3 + t, 128);
      if (qd == 0 && lane == 0) {
        tma_store_3d(&P.tmO, stg, h * kAttnD, qt * kAttnItemQ + t * kAttnBQ, b);
        tma_store_commit();
      }
      if (P.lse && qi < P.Lq)
        P.lse[(static_cast<size_t>(b) * P.H + h) * P.Lq + qi] = (l_run > 0.f) ? fmaf(m_run, 0.69314718055994531f, __logf(l_run)) : -INFINITY;
    }
    if (pingpong && t == 0) named_bar_sync(1, 256);
    if (qd == 0 && lane == 0) tma_store_wait<0>();           // every output tile written before the CTA retires
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace fnd
