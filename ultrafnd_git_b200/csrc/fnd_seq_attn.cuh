// fnd_seq_attn.cuh — multi-head cross-attention forward for the sequence front-end (Tier B), flash-style on tcgen05.
//
//   O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] * scale + key_padding_mask[b,k] ) V[b,k,h,:]       d_k = 64
//
// One CTA owns one (sample, head, 128-query tile) and walks the key/value sequence in blocks of 64:
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (SWIZZLE_128B, 3-D maps: rows past the sequence end of
//               THIS sample are zero-filled) through a 3-stage ring
//   warp 1      one elected thread issues tcgen05.mma:  S_j = Q K_j^T (128x64x64, into one of two TMEM S buffers — S_{j+1}
//               is issued BEFORE P_j V_j so the tensor pipe works while the softmax warps are busy) and
//               T_j = P_j V_j (128x64x64, P from shared memory, V MN-major, into one of two TMEM buffers)
//   warps 2..5  softmax: thread r owns query row r — tcgen05.ld of its S row, key-padding mask, running max / sum
//               (online softmax, exp2 with the scale folded in), P_j as bf16 into swizzled shared memory for the next
//               MMA, O accumulated in REGISTERS (acc = acc * alpha_j + T_{j-1}: no TMEM read-modify-write correction pass)
// TMEM: 2 x 64 (S) + 2 x 64 (T) = 256 columns and ~98 KB of shared memory per CTA -> two CTAs per SM, whose MMA / softmax
// phases interleave on the SM. A query row with no valid key yields zeros (LSE = -inf).
//
// No counterpart in the reference (the reference's "co-attention" is a per-sample sigmoid gate,
// src/models/fusion/cross_modal_transformer.py:39-55); checked against the self-oracle oracle/seq_oracle.py.
#pragma once
#include "fnd_common.cuh"

namespace fnd {

constexpr int kAttnBQ = 128;                 // query rows per CTA
constexpr int kAttnBK = 64;                  // keys per block
constexpr int kAttnD = 64;                   // head dimension
constexpr int kAttnStages = 3;
constexpr int kAttnThreads = 192;
constexpr int kAttnTmemCols = 256;
constexpr int kAttnQBytes = kAttnBQ * kAttnD * 2;        // 16 KB
constexpr int kAttnPBytes = kAttnBQ * kAttnBK * 2;       // 16 KB
constexpr int kAttnKBytes = kAttnBK * kAttnD * 2;        // 8 KB
constexpr int kAttnSmemBytes = 1024 /*align*/ + 1024 /*barriers*/ + kAttnQBytes + 2 * kAttnPBytes + kAttnStages * 2 * kAttnKBytes;

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
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2) seq_attn_fwd_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = q_full + 1;
  uint64_t* kv_empty = kv_full + kAttnStages;
  uint64_t* s_full = kv_empty + kAttnStages;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  uint8_t* sQ = smem + 1024;
  uint8_t* sP = sQ + kAttnQBytes;
  uint8_t* sKV = sP + 2 * kAttnPBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kAttnBQ;
  int kv_len = P.Lk;
  if (P.kv_len) kv_len = min(max(P.kv_len[b], 0), P.Lk);
  const int nblk = (kv_len + kAttnBK - 1) / kAttnBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < kAttnStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 128);
        mbar_init(&p_full[i], 128);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 128);
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
    if (lane == 0 && nblk > 0) {
      mbar_arrive_expect_tx(q_full, kAttnQBytes);
      tma_load_3d(sQ, &P.tmQ, q_full, P.q_col0 + h * kAttnD, q0, b, kEvictFirst);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = static_cast<uint32_t>(j / kAttnStages) & 1u;
        if (!mbar_wait(&kv_empty[s], ph ^ 1u, P.err, FND_DEV_TIMEOUT_PRODUCER)) break;
        mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnKBytes);
        uint8_t* sK = sKV + s * 2 * kAttnKBytes;
        tma_load_3d(sK, &P.tmK, &kv_full[s], P.k_col0 + h * kAttnD, j * kAttnBK, b, kEvictLast);
        tma_load_3d(sK + kAttnKBytes, &P.tmV, &kv_full[s], P.v_col0 + h * kAttnD, j * kAttnBK, b, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && nblk > 0) {
      const uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBK, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnD, 0, 1);       // T = P V   : V is [keys][d] = MN-major B
      const uint32_t qBase = smem_u32(sQ);
      bool ok = mbar_wait(q_full, 0u, P.err, FND_DEV_TIMEOUT_MMA);
      auto issue_s = [&](int jj) {
        const int s = jj % kAttnStages;
        ok = ok && mbar_wait(&kv_full[s], static_cast<uint32_t>(jj / kAttnStages) & 1u, P.err, FND_DEV_TIMEOUT_MMA);
        ok = ok && mbar_wait(&s_free[jj & 1], (static_cast<uint32_t>(jj >> 1) & 1u) ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) return;
        tc_fence_after_sync();
        const uint32_t kBase = smem_u32(sKV + s * 2 * kAttnKBytes);
        const uint32_t tS = tmem_base + static_cast<uint32_t>((jj & 1) * kAttnBK);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_f16(tS, make_smem_desc_sw128(qBase + k * 32, 16, 1024), make_smem_desc_sw128(kBase + k * 32, 16, 1024),
                   idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[jj & 1]);
      };
      issue_s(0);
      for (int j = 0; j < nblk && ok; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const uint32_t par = static_cast<uint32_t>(j >> 1) & 1u;
        ok = ok && mbar_wait(&p_full[j & 1], par, P.err, FND_DEV_TIMEOUT_MMA);
        ok = ok && mbar_wait(&o_free[j & 1], par ^ 1u, P.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        tc_fence_after_sync();
        const int s = j % kAttnStages;
        const uint32_t pBase = smem_u32(sP + (j & 1) * kAttnPBytes);
        const uint32_t vBase = smem_u32(sKV + s * 2 * kAttnKBytes + kAttnKBytes);
        const uint32_t tO = tmem_base + static_cast<uint32_t>(2 * kAttnBK + (j & 1) * kAttnD);
#pragma unroll
        for (int k = 0; k < kAttnBK / 16; ++k)
          umma_f16(tO, make_smem_desc_sw128(pBase + k * 32, 16, 1024), make_smem_desc_sw128(vBase + k * 2048, 8192, 1024),
                   idesc_o, k != 0 ? 1u : 0u);
        umma_commit(&o_full[j & 1]);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ================= softmax + output: warps 2..5, thread = one query row =================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int qi = q0 + row;
    // off_run = running row maximum in exp2 units (m * scale * log2 e), -inf while no valid key has been seen
    float off_run = -INFINITY, l_run = 0.f;
    float acc[kAttnD];
#pragma unroll
    for (int i = 0; i < kAttnD; ++i) acc[i] = 0.f;
    const unsigned char* mrow = P.kv_mask ? P.kv_mask + static_cast<size_t>(b) * P.Lk : nullptr;
    bool ok = true;

    auto add_tmp = [&](int jj) {     // acc += T_jj  (the P_jj V_jj product), then release the TMEM buffer
      ok = ok && mbar_wait(&o_full[jj & 1], static_cast<uint32_t>(jj >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      const uint32_t tO = tmem_base + lane_addr + static_cast<uint32_t>(2 * kAttnBK + (jj & 1) * kAttnD);
#pragma unroll
      for (int c = 0; c < kAttnD; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tO + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c + i] += __uint_as_float(r[i]);
      }
      tc_fence_before_sync();
      mbar_arrive(&o_free[jj & 1]);
    };

#pragma unroll 1
    for (int j = 0; j < nblk; ++j) {
      ok = ok && mbar_wait(&s_full[j & 1], static_cast<uint32_t>(j >> 1) & 1u, P.err, FND_DEV_TIMEOUT_EPILOGUE);
      tc_fence_after_sync();
      float s[kAttnBK];
      const uint32_t tS = tmem_base + lane_addr + static_cast<uint32_t>((j & 1) * kAttnBK);
#pragma unroll
      for (int c = 0; c < kAttnBK; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tS + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c + i] = __uint_as_float(r[i]);
      }
      tc_fence_before_sync();
      mbar_arrive(&s_free[j & 1]);

      // ---- key-padding mask of this block as a 64-bit validity word (warp-cooperative: two ballots) ----
      const int k0 = j * kAttnBK;
      unsigned long long valid = ~0ull;
      if (mrow || k0 + kAttnBK > kv_len) {
        const int ka = k0 + lane, kb = k0 + 32 + lane;
        const bool va = ka < kv_len && (!mrow || mrow[ka] != 0);
        const bool vb = kb < kv_len && (!mrow || mrow[kb] != 0);
        valid = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, va)) |
                (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, vb)) << 32);
      }
      float mx = -INFINITY;
      if (valid == ~0ull) {
#pragma unroll
        for (int i = 0; i < kAttnBK; ++i) mx = fmaxf(mx, s[i]);
      } else {
#pragma unroll
        for (int i = 0; i < kAttnBK; ++i) {
          s[i] = ((valid >> i) & 1ull) ? s[i] : -INFINITY;
          mx = fmaxf(mx, s[i]);
        }
      }
      const float off_new = fmaxf(off_run, mx * P.scale_log2);              // scale_log2 > 0: max commutes with the scaling
      const float off = (off_new == -INFINITY) ? 0.f : off_new;
      const float alpha = ex2_approx(off_run - off);                        // off_run = -inf -> 0; unchanged maximum -> exactly 1
      float psum = 0.f;
      uint32_t pk[kAttnBK / 2];
#pragma unroll
      for (int i = 0; i < kAttnBK; i += 2) {
        const float p0 = ex2_approx(fmaf(s[i], P.scale_log2, -off));
        const float p1 = ex2_approx(fmaf(s[i + 1], P.scale_log2, -off));
        psum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      l_run = fmaf(l_run, alpha, psum);
      off_run = off_new;
      // ---- P_j -> shared memory in the SWIZZLE_128B K-major layout the MMA descriptor expects:
      //      row r at r*128 B, 16-byte chunk c stored at chunk (c ^ (r & 7)) ----
      {
        uint8_t* prow = sP + (j & 1) * kAttnPBytes + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&p_full[j & 1]);
      // ---- fold in the previous block's product while the tensor pipe works on this one ----
      if (j > 0) add_tmp(j - 1);
      if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int i = 0; i < kAttnD; ++i) acc[i] *= alpha;
      }
    }
    if (nblk > 0) add_tmp(nblk - 1);

    if (qi < P.Lq) {
      const float inv = (ok && l_run > 0.f) ? __fdividef(1.f, l_run) : 0.f;
      __nv_bfloat16* op = P.out + (static_cast<size_t>(b) * P.Lq + qi) * P.out_pitch + h * kAttnD;
#pragma unroll
      for (int c = 0; c < kAttnD; c += 8)
        *reinterpret_cast<uint4*>(op + c) = make_uint4(pack_bf16x2(acc[c] * inv, acc[c + 1] * inv), pack_bf16x2(acc[c + 2] * inv, acc[c + 3] * inv),
                                                       pack_bf16x2(acc[c + 4] * inv, acc[c + 5] * inv), pack_bf16x2(acc[c + 6] * inv, acc[c + 7] * inv));
      if (P.lse)
        P.lse[(static_cast<size_t>(b) * P.H + h) * P.Lq + qi] = (l_run > 0.f) ? fmaf(off_run, 0.69314718055994531f, __logf(l_run)) : -INFINITY;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace fnd
