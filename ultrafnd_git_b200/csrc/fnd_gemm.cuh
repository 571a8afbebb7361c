// fnd_gemm.cuh — grouped, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]      (bf16 operands, fp32 accumulation in TMEM)
//
// One launch runs a TABLE of independent problems (GemmProblem, passed by value in kernel-parameter space); each CTA
// owns one (tile_m, tile_n, k-split) work item of one problem. Operands arrive by TMA (SWIZZLE_128B) into a deep
// shared-memory ring (as many stages as fit in ~208 KB: 6..11), a single elected thread issues tcgen05.mma
// (UMMA 128 x BN x 16, BN = 16/32/64/128), the accumulator lives in TMEM and EIGHT epilogue warps (two per TMEM lane
// quarter, splitting the column groups) read it back with tcgen05.ld.
//
// Why this shape: at the reference's batch (128) every GEMM of the step is a latency chain, not a throughput problem
// (DESIGN.md §4). The ring is deep enough that ALL k-blocks of a K<=512 problem are in flight at once; narrow tiles
// put 16..144 CTAs on the 148 SMs instead of 8; two epilogue warps per scheduler hide each other's dependent-issue
// latency; and the kernel is PDL-aware: it becomes resident under its predecessor, sets up barriers/TMEM and already
// streams its WEIGHT tiles (static data) before griddepcontrol.wait releases the activation loads.
//
// Either operand may be K-major (contraction index contiguous in memory: activations x weights of nn.Linear) or
// MN-major (row index contiguous: what dgrad needs for W and wgrad needs for dY and X), selected by launch-uniform
// flags, so forward, dgrad and wgrad all read the SAME row-major tensors — no transposed copies are ever materialised.
//
// Precision modes: ncombo == 1 is plain bf16; ncombo == 3 is "fp32x3": every operand is stored as a bf16 (hi, lo)
// pair and the k-loop runs the three products Ah*Bh + Ah*Bl + Al*Bh, which reproduces fp32 GEMM to ~2^-17 relative
// while staying on the tensor pipe.
//
// Split-K: the `splits` CTAs of one output tile are co-resident (the host keeps split launches <= 148 CTAs, one CTA
// per SM). Each writes its fp32 partial tile to an L2-resident workspace and arrives on a per-tile counter; once all
// have arrived EVERY CTA of the tile finishes a disjoint share of the tile (32-row x 8-column units, round-robin over
// the splits), summing the partials in fixed split order (deterministic) and running the epilogue on its share — the
// fix-up is spread over all splits instead of serialising on the last arriver.
//
// Replaces (reference, all via ATen addmm on CPU): nn.Linear forward/backward at
//   src/models/fusion/cross_modal_transformer.py:96-102,122-129,147-150,186,197 and
//   src/models/fusion/deep_truth_classifier.py:121-128,162.
#pragma once
#include "fnd_common.cuh"
#include "fnd_rows.cuh"

namespace fnd {

constexpr int kGemmBM = 128;          // UMMA M (rows of A per tile)
constexpr int kGemmBK = 64;           // contraction elements per stage (128 B of bf16)
constexpr int kGemmMaxStages = 12;
constexpr int kGemmStageBytesA = kGemmBM * kGemmBK * 2;   // 16 KB
constexpr int kGemmOperandBudget = 208 * 1024;            // operand ring (upper bound; a launch asks for what it uses)
constexpr int kGemmSmemHeader = 1024;                     // barriers, TMEM slot, reduction scratch (in front of the ring)
constexpr int kGemmSmemBytes = kGemmOperandBudget + 1024 /*align slack*/ + kGemmSmemHeader;
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmEpiThreads = kGemmEpiWarps * 32;
constexpr int kGemmThreads = 64 + kGemmEpiThreads;        // warp0 TMA, warp1 MMA+TMEM, warps2-9 epilogue
constexpr int kGemmTmemCols = 128;

// Everything the epilogue may do to an accumulator value v at (m, n), in this order.
struct EpiParams {
  const float* bias;                 // v += bias[n]
  const float* aux;                  // rank-2 update: v += aux[m,0]*aux_w[n,0] + aux[m,1]*aux_w[n,1]
  const float* aux_w;
  int aux_w_pitch;
  const float* add_in;               // v += add_in[m*add_pitch + n]
  int add_pitch;
  float* out_pre;                    // store v (pre-activation) for the backward pass
  int pre_pitch;
  int act;                           // 1: v = gelu_erf(v)
  float drop_p;                      // >0: v *= dropout_mask(drop_stream, m*N+n) / (1-p)
  int drop_stream;
  const float* gate_z;               // backward gate: v *= gelu'(gate_z[m,n]) * dropout_mask(gate_stream)/(1-gate_p)
  int gate_pitch;
  float gate_p;
  int gate_stream;
  float* out_f32;                    // store v as fp32
  int f32_pitch;
  __nv_bfloat16* out_hi;             // store bf16(v) (and the residual in out_lo when non-null)
  __nv_bfloat16* out_lo;
  int bf_pitch;
  float* sumsq_slots;                // per-CTA sum of v^2 (for the global gradient norm)
  int plain_f32;                     // host-set: the epilogue is ONLY "store v as fp32 (+ sum of squares)", pitch % 8 == 0;
                                     // 2 = additionally the operand ring is large enough to stage the tile for coalesced stores
  unsigned int* done_ctr;            // non-null: bumped once per tile after its stores (consumers in the same launch wait on it)
  __nv_bfloat16* out_bf;             // store-only epilogues: also store bf16(v) at the same [m * f32_pitch + n] offsets (the
                                     // bf16 gradient mirror that the data-parallel step reduces through the NVSwitch)
  int routed;                        // host-set: out_f32 lies in the gradient arena (pitch % 8 == 0); with an active DpRoute the
                                     // store-only epilogue sends each row segment straight to its OWNER's staging slot
};

// Data-parallel routing of weight-gradient tiles (fused reduce-scatter push, bf16 on the wire). The hot arena [0, n_hot)
// is cut into three ranges (0: [a0, a1), 1: [0, a0), 2: [a1, n_hot)); rank p owns elements [p * per[s], (p+1) * per[s]) of
// range s (per[s] a multiple of 1024) and keeps, for every source rank q, a staging slot of slot_cap bf16 elements in
// which its piece of range s starts at goff[p][s]. A 64-element row segment of a weight gradient (64-aligned in the
// arena) therefore has exactly one owner. world == 0: routing off.
struct DpRoute {
  int rank, world;
  const float* grads;                // this rank's gradient arena (arena index = pointer difference)
  __nv_bfloat16* stage[8];           // staging buffer of every rank (peer-mapped)
  uint32_t slot_cap;
  uint32_t a0, a1;
  uint32_t r_lo[3], per[3];
  uint32_t goff[8][3];
};
__device__ __forceinline__ __nv_bfloat16* dp_route_dst(const DpRoute& R, uint32_t i) {
  const int s = (i >= R.a0) ? (i < R.a1 ? 0 : 2) : 1;
  const uint32_t rel = i - R.r_lo[s];
  const uint32_t owner = rel / R.per[s];
  return R.stage[owner] + (static_cast<size_t>(R.rank) * R.slot_cap + R.goff[owner][s] + (rel - owner * R.per[s]));
}

struct alignas(128) GemmProblem {
  CUtensorMap tmA[2];                // [0] hi, [1] lo
  CUtensorMap tmB[2];
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total;
  int cta_begin, cta_count;
  int ncombo;                        // 1 bf16, 3 fp32x3
  int bn;                            // tile width: 16/32/64/128 (MN-major B: 64/128)
  int nstages, stage_bytes;          // operand ring geometry (host: fill_problem)
  int b_static;                      // 1: operand B is static data (weights): its loads may precede griddepcontrol.wait
  unsigned long long hintA, hintB;   // L2 eviction hints for the two operand streams
  float* splitk_ws;                  // [tiles][splits][bn/8][128][8] fp32
  int* splitk_ctr;                   // [tiles][2] (arrive, depart), zero between launches
  EpiParams epi;
};

struct RunCtx {
  int* err;                          // device error flag
  const uint32_t* rng;               // [0] seed lo, [1] seed hi, [2] fusion salt, [3] classifier salt
  int training;                      // 0: every dropout (forward masks and backward gates) is the identity
  long long* dbg;                    // optional [grid][8] clock64 stamps (probe builds only; null in production)
  DpRoute route;                     // data-parallel fused push (weight-gradient launch only; world == 0 otherwise)
};
#define FND_STAMP(i) do { if (ctx.dbg) ctx.dbg[static_cast<size_t>(blockIdx.x) * 8 + (i)] = clock64(); } while (0)

__device__ __forceinline__ void epi_named_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// The problem table travels BY VALUE in kernel-parameter space (constant bank): no global-memory fetch of table or
// TMA descriptors sits on the critical path of a launch (after an L2 flush those were two dependent HBM round trips).
// ONE kernel binary serves forward, dgrad and wgrad (operand majorness is a launch-uniform runtime flag): a training
// step issues 12 GEMM launches, and sharing the code keeps it warm in the instruction caches between them.
constexpr int kGemmTableCap = 16;
struct GemmTableP {
  GemmProblem p[kGemmTableCap];
  int nprob;
  int a_mn, b_mn;     // 0: K-major, 1: MN-major
  int gemm_ctas;      // CTAs [gemm_ctas, gridDim.x) of a kVariant == 1 launch run finalize jobs instead of a tile
  int cluster_k;      // 1: the launch's clusters are the k-splits of a tile (cluster size == splits of every problem): the
                      //    split-K partials are exchanged through distributed shared memory instead of an L2 workspace
};

struct EpiCtx {
  DropCfg dfw, dbw;
  uint32_t key_fw, key_bw;
};

// One 8-column group of one output row: everything that happens after the accumulator value is known.
__device__ __forceinline__ float epi_group(const EpiParams& E, const EpiCtx& X, float (&v)[8], int m, int n0, int PN,
                                           float a0, float a1) {
  if (E.bias) {
    const float4 b0 = ldg_f4(E.bias + n0), b1 = ldg_f4(E.bias + n0 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (E.aux) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 w = __ldg(reinterpret_cast<const float2*>(E.aux_w + static_cast<size_t>(n0 + j) * E.aux_w_pitch));
      v[j] = fmaf(a0, w.x, fmaf(a1, w.y, v[j]));
    }
  }
  // pitches of add_in / out_pre / gate_z are multiples of 8 floats (workspace buffers): 256-bit accesses
  if (E.add_in) {
    float t[8];
    ldg_f8(E.add_in + static_cast<size_t>(m) * E.add_pitch + n0, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t[j];
  }
  if (E.out_pre) st_f8(E.out_pre + static_cast<size_t>(m) * E.pre_pitch + n0, v);
  if (E.act == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
  }
  const uint64_t e0 = static_cast<uint64_t>(m) * static_cast<uint64_t>(PN) + n0;   // multiple of 8
  if (X.dfw.p > 0.f) {
    float mm[8];
    dropout_mult8(X.dfw, X.key_fw, e0 >> 3, mm);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= mm[j];
  }
  if (E.gate_z) {
    float z[8];
    ldg_f8(E.gate_z + static_cast<size_t>(m) * E.gate_pitch + n0, z);
    float mm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (X.dbw.p > 0.f) dropout_mult8(X.dbw, X.key_bw, e0 >> 3, mm);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= gelu_erf_grad(z[j]) * mm[j];
  }
  if (E.out_f32) {
    float* dst = E.out_f32 + static_cast<size_t>(m) * E.f32_pitch + n0;
    if ((E.f32_pitch & 7) == 0) {
      st_f8(dst, v);
    } else if ((E.f32_pitch & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {                                   // even pitch (pre.0.weight: 514): 8-byte aligned rows
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
    }
  }
  if (E.out_hi) {
    const size_t o = static_cast<size_t>(m) * E.bf_pitch + n0;
    uint32_t ph[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ph[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(E.out_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (E.out_lo) {
      uint32_t pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&ph[j]);
        pl[j] = pack_bf16x2(v[2 * j] - __low2float(h2), v[2 * j + 1] - __high2float(h2));
      }
      *reinterpret_cast<uint4*>(E.out_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
  float ss = 0.f;
  if (E.sumsq_slots) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
  }
  return ss;
}

// kVariant 0: general kernel (one CTA per SM, split-K capable).
// kVariant 1: "light" kernel for the weight-gradient launch: no split-K, register-capped so that TWO CTAs share an SM
//             (one CTA's store-bound epilogue overlaps the other's loads and MMAs; its ring is only as deep as its
//             k-loop), and the trailing CTAs [gemm_ctas, gridDim.x) run the finalize jobs (bias / threshold / leaf /
//             evidence gradient reductions, fnd_rows.cuh) concurrently with the tiles instead of in a kernel of their own.
template <int kVariant>
__global__ void __launch_bounds__(kGemmThreads, kVariant == 1 ? 2 : 1)
fnd_gemm_kernel(const __grid_constant__ GemmTableP tbl, RunCtx ctx, const __grid_constant__ FinParams fin) {
  if (kVariant == 1 && static_cast<int>(blockIdx.x) >= tbl.gemm_ctas) {
    griddep_wait();
    griddep_launch();
    if (threadIdx.x < 256) finalize_cta(fin, static_cast<int>(blockIdx.x) - tbl.gemm_ctas, static_cast<int>(gridDim.x));
    return;
  }
  const GemmProblem* probs = tbl.p;
  const int nprob = tbl.nprob;
  const bool A_MN = tbl.a_mn != 0, B_MN = tbl.b_mn != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kGemmMaxStages;
  uint64_t* accum_bar = empty_bar + kGemmMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* red_smem = reinterpret_cast<float*>(tmem_slot + 2);   // 8 floats
  smem += kGemmSmemHeader;                                     // operand ring

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FND_STAMP(0);

  // ---- locate this CTA's work item ----
  int pi = 0;
  while (pi + 1 < nprob && static_cast<int>(blockIdx.x) >= probs[pi + 1].cta_begin) ++pi;
  const GemmProblem& P = probs[pi];
  const int local = static_cast<int>(blockIdx.x) - P.cta_begin;
  const int splits = P.splits;
  const int split = local % splits;
  const int tile = local / splits;
  const int tm = tile % P.tiles_m;
  const int tn = tile / P.tiles_m;
  const int bn = P.bn;
  const int ncombo = P.ncombo;
  const int nstages = P.nstages;
  const int stage_bytes = P.stage_bytes;
  const int kb0 = split * P.kb_per_split;
  const int kb1 = min(kb0 + P.kb_per_split, P.kb_total);
  const int iters = (kb1 - kb0) * ncombo;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA[0]);
    tma_prefetch_desc(&P.tmB[0]);
    if (ncombo > 1) {
      tma_prefetch_desc(&P.tmA[1]);
      tma_prefetch_desc(&P.tmB[1]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nstages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kGemmTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) FND_STAMP(1);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t bytesB = static_cast<uint32_t>(bn) * kGemmBK * 2;
      const uint32_t tx = kGemmStageBytesA + bytesB;
      auto load_a = [&](int it) {
        const int s = it % nstages;
        const int kb = kb0 + it / ncombo;
        const int c = it - (it / ncombo) * ncombo;
        const void* mapA = &P.tmA[c == 2 ? 1 : 0];
        uint8_t* sA = smem + s * stage_bytes;
        if (!A_MN) {
          tma_load_2d(sA, mapA, &full_bar[s], kb * kGemmBK, tm * kGemmBM, P.hintA);
        } else {
          tma_load_2d(sA, mapA, &full_bar[s], tm * kGemmBM, kb * kGemmBK, P.hintA);
          tma_load_2d(sA + 8192, mapA, &full_bar[s], tm * kGemmBM + 64, kb * kGemmBK, P.hintA);
        }
      };
      auto load_b = [&](int it) {
        const int s = it % nstages;
        const int kb = kb0 + it / ncombo;
        const int c = it - (it / ncombo) * ncombo;
        const void* mapB = &P.tmB[c == 1 ? 1 : 0];
        uint8_t* sB = smem + s * stage_bytes + kGemmStageBytesA;
        if (!B_MN) {
          tma_load_2d(sB, mapB, &full_bar[s], kb * kGemmBK, tn * bn, P.hintB);
        } else {
          for (int ch = 0; ch < bn / 64; ++ch)
            tma_load_2d(sB + ch * 8192, mapB, &full_bar[s], tn * bn + ch * 64, kb * kGemmBK, P.hintB);
        }
      };
      // Static operand (weights): fill the ring's first pass before waiting for the predecessor kernel.
      const int npre = P.b_static ? min(iters, nstages) : 0;
      for (int it = 0; it < npre; ++it) {
        mbar_arrive_expect_tx(&full_bar[it], tx);
        load_b(it);
      }
      griddep_wait();
      griddep_launch();
      for (int it = 0; it < npre; ++it) load_a(it);
      for (int it = npre; it < iters; ++it) {
        const int s = it % nstages;
        const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
        if (!mbar_wait(&empty_bar[s], ph ^ 1u, ctx.err, FND_DEV_TIMEOUT_PRODUCER)) break;
        mbar_arrive_expect_tx(&full_bar[s], tx);
        load_a(it);
        load_b(it);
      }
    }
    __syncwarp();
    if (kVariant == 0 && tbl.cluster_k && splits > 1) { cluster_arrive(); cluster_wait(); cluster_arrive(); cluster_wait(); }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kGemmBM, bn, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t kstepA = A_MN ? 2048u : 32u, kstepB = B_MN ? 2048u : 32u;
      const uint32_t lboA = A_MN ? 8192u : 16u, lboB = B_MN ? 8192u : 16u;
      bool ok = true;
      for (int it = 0; it < iters && ok; ++it) {
        const int s = it % nstages;
        const uint32_t ph = static_cast<uint32_t>(it / nstages) & 1u;
        ok = mbar_wait(&full_bar[s], ph, ctx.err, FND_DEV_TIMEOUT_MMA);
        if (!ok) break;
        if (it == 0) FND_STAMP(2);
        tc_fence_after_sync();
        const uint32_t aBase = smem_u32(smem + s * stage_bytes);
        const uint32_t bBase = aBase + kGemmStageBytesA;
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          // K-major: 16 contraction elements = 32 B inside the 128-B swizzled row.
          // MN-major: 16 contraction rows of 128 B = 2048 B (two 8-row swizzle atoms).
          const uint64_t ad = make_smem_desc_sw128(aBase + k * kstepA, lboA, 1024);
          const uint64_t bd = make_smem_desc_sw128(bBase + k * kstepB, lboB, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);      // frees the smem stage once these MMAs retire
      }
      umma_commit(accum_bar);            // accumulator complete
      FND_STAMP(3);
    }
    __syncwarp();
    if (kVariant == 0 && tbl.cluster_k && splits > 1) { cluster_arrive(); cluster_wait(); cluster_arrive(); cluster_wait(); }
  } else {
    // ================= epilogue (warps 2..9) =================
    // Deliberately ROLLED (8 columns per iteration, #pragma unroll 1): a fully unrolled epilogue was ~100 KB of
    // straight-line SASS and made every small launch instruction-fetch bound (~25 us fixed cost, measured).
    // By VALUE: the inline-asm tcgen05/mbarrier wrappers carry "memory" clobbers, so fields read through a pointer
    // into the problem table were re-loaded after every one of them. Registers are plentiful here.
    const EpiParams E = P.epi;
    const int PM = P.M, PN = P.N;
    float* const ws_base = P.splitk_ws;
    int* const ctr_base = P.splitk_ctr;
    const int ew = warp - 2;                       // 0..7
    const int lane_grp = warp & 3;                 // TMEM lane quarter this warp may access (hardware rule: warp % 4)
    const int half = ew >> 2;                      // the two warps of a lane quarter alternate column groups
    const int row = lane_grp * 32 + lane;
    const int m = tm * kGemmBM + row;
    const bool row_ok = m < PM;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const int ngroups = bn / 8;
    const int epi_tid = threadIdx.x - 64;
    const int nb = tn * bn;
    // static epilogue operand: pull the tile's bias slice towards L2 before the predecessor finishes
    if (E.bias && epi_tid * 32 < bn) prefetch_l2(E.bias + nb + epi_tid * 32);
    griddep_wait();
    // While the main loop runs, pull this thread's row of the activation-side epilogue operands towards L2 (they are
    // cold after the optimizer has streamed ~0.5 GB through the cache).
    if (row_ok && half == 0) {
      if (E.gate_z)
        for (int c = 0; c < bn; c += 32) prefetch_l2(E.gate_z + static_cast<size_t>(m) * E.gate_pitch + nb + c);
      if (E.add_in)
        for (int c = 0; c < bn; c += 32) prefetch_l2(E.add_in + static_cast<size_t>(m) * E.add_pitch + nb + c);
    }
    const uint64_t seed = ctx.rng ? ((static_cast<uint64_t>(ctx.rng[1]) << 32) | ctx.rng[0]) : 0ull;
    EpiCtx X;
    X.dfw = make_dropcfg(ctx.training ? E.drop_p : 0.f, seed);
    X.dbw = make_dropcfg(ctx.training ? E.gate_p : 0.f, seed);
    X.key_fw = stream_key(ctx.rng, E.drop_stream);
    X.key_bw = stream_key(ctx.rng, E.gate_stream);

    const bool proceed = mbar_wait(accum_bar, 0u, ctx.err, FND_DEV_TIMEOUT_EPILOGUE);
    tc_fence_after_sync();
    if (epi_tid == 0) FND_STAMP(4);
    float ss = 0.f;

    if (kVariant == 1 || splits == 1) {
      float a0 = 0.f, a1 = 0.f;
      if (E.aux && row_ok) {
        a0 = E.aux[static_cast<size_t>(m) * 2];
        a1 = E.aux[static_cast<size_t>(m) * 2 + 1];
      }
      if (epi_tid == 0) FND_STAMP(5);
      if (kVariant == 1 && E.routed && ctx.route.world > 1 && E.plain_f32 == 2 && bn >= 64) {
        // Data-parallel fused push: the reduce-scatter's data movement happens HERE. Each warp packs its 32 rows x 64
        // columns of the accumulator to bf16 through shared memory and stores every row segment (128 contiguous bytes,
        // 8 lanes x 16 B) into the staging slot of the rank that OWNS that arena range — over NVLink for 7 of 8 segments
        // on 8 GPUs, locally for its own. The fp32 gradient is never written to HBM and no separate push kernel re-reads it.
        constexpr int kRowB = 144;                                       // staged row: 64 bf16 + 16 B pad (conflict-free)
        uint8_t* stg = smem + ew * (32 * kRowB);
        const uint32_t base_idx = static_cast<uint32_t>(E.out_f32 - ctx.route.grads);
        const int cw = bn >> 1;                                          // columns per warp: 64 (bn 128) or 32 (bn 64)
        const int rsub = lane >> 3, ch8 = lane & 7;
#pragma unroll 1
        for (int c = half * cw; c < (half + 1) * cw; c += 64) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh * 32 < cw) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c + hh * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(stg + lane * kRowB + hh * 64 + q * 16) =
                    make_uint4(pack_bf16x2(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                               pack_bf16x2(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                               pack_bf16x2(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                               pack_bf16x2(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
            }
          }
          __syncwarp();
          const int n = nb + c + ch8 * 8;
          if (proceed && n < PN && ch8 * 8 < cw) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = k * 4 + rsub;
              const int mm = tm * kGemmBM + lane_grp * 32 + rr;
              if (mm < PM) {
                const uint4 v4 = *reinterpret_cast<const uint4*>(stg + rr * kRowB + ch8 * 16);
                __nv_bfloat16* dst = dp_route_dst(ctx.route, base_idx + static_cast<uint32_t>(mm) * E.f32_pitch + nb + c);
                *reinterpret_cast<uint4*>(dst + ch8 * 8) = v4;
              }
            }
          }
          __syncwarp();
        }
      } else if (E.plain_f32 == 2 && bn >= 64) {
        // Store-only epilogue (weight gradients, dcat), coalesced: a thread owns a ROW of the accumulator, so storing
        // straight from registers makes every warp-level store touch 32 different 128-byte lines (two LSU wavefronts
        // per lane: the 780-tile wgrad launch was LSU-bound, ncu). Instead each warp transposes its 32x32 block through
        // shared memory (the operand ring is dead once the accumulator is complete) and writes 4 full rows x 128 B
        // per instruction.
        constexpr int kStgPitch = 36;                                  // floats per staged row (16-byte aligned, conflict-free)
        float* stg = reinterpret_cast<float*>(smem) + ew * (32 * kStgPitch);
        const int c_end = (half + 1) * (bn >> 1);
        const int rsub = lane >> 3, csub = (lane & 7) * 4;
#pragma unroll 1
        for (int c = half * (bn >> 1); c < c_end; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v4 = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                          __uint_as_float(r[4 * q + 3]));
            *reinterpret_cast<float4*>(stg + lane * kStgPitch + 4 * q) = v4;
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
              if (mm < PM) {
                const float4 v4 = *reinterpret_cast<const float4*>(stg + rr * kStgPitch + csub);
                const size_t o = static_cast<size_t>(mm) * E.f32_pitch + n;
                *reinterpret_cast<float4*>(E.out_f32 + o) = v4;
                if (E.out_bf) *reinterpret_cast<uint2*>(E.out_bf + o) = make_uint2(pack_bf16x2(v4.x, v4.y), pack_bf16x2(v4.z, v4.w));
              }
            }
          }
          __syncwarp();
        }
      } else if (E.plain_f32 && bn >= 64) {
        // Store-only epilogue, direct: 32 accumulator columns per tcgen05.ld, four 256-bit stores per thread.
        // Each warp of a lane quarter takes one contiguous half of the tile's columns.
        const int c_end = (half + 1) * (bn >> 1);
#pragma unroll 1
        for (int c = half * (bn >> 1); c < c_end; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          if (!proceed || !row_ok) continue;
          float* dst = E.out_f32 + static_cast<size_t>(m) * E.f32_pitch + nb + c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (nb + c + q * 8 < PN) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
              st_f8(dst + q * 8, v);
              if (E.out_bf)
                *reinterpret_cast<uint4*>(E.out_bf + (dst - E.out_f32) + q * 8) =
                    make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
#pragma unroll
              for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
            }
          }
        }
      } else
#pragma unroll 1
      for (int g = half; g < ngroups; g += 2) {
        uint32_t r[8];
        tmem_ld_32x8(taddr + g * 8, r);
        tmem_ld_wait();
        const int n0 = nb + g * 8;
        if (!proceed || !row_ok || n0 >= PN) continue;           // N is a multiple of 8 (checked on the host)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
        ss += epi_group(E, X, v, m, n0, PN, a0, a1);
      }
    } else if (kVariant == 0 && tbl.cluster_k) {
      // ---- split-K inside a cluster: the `splits` CTAs of a tile ARE the cluster (rank == split). Each CTA parks its partial
      // tile in its own shared memory (the operand ring is dead: every MMA has retired), one cluster barrier replaces the
      // fence + atomic-counter rendezvous, and every split sums its share of the tile straight out of its peers' shared
      // memory (ld.shared::cluster) in the same fixed split order as the L2 path below — bit-identical results, no
      // workspace round trip through L2. A second cluster barrier keeps every CTA's partial alive until its peers have read it.
      float* const part = reinterpret_cast<float*>(smem);            // [group][row][8]
#pragma unroll 1
      for (int g = half; g < ngroups; g += 2) {
        uint32_t r[8];
        tmem_ld_32x8(taddr + g * 8, r);
        tmem_ld_wait();
        float* dst = part + (static_cast<size_t>(g) * kGemmBM + row) * 8;
        *reinterpret_cast<uint4*>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<uint4*>(dst + 4) = make_uint4(r[4], r[5], r[6], r[7]);
      }
      cluster_arrive();
      cluster_wait();
      if (epi_tid == 0) FND_STAMP(5);
      const int nunits = ngroups * 4;
      const uint32_t part_u32 = smem_u32(part);
#pragma unroll 1
      for (int u = split + ew * splits; u < nunits; u += kGemmEpiWarps * splits) {
        const int g = u >> 2, q = u & 3;
        const int urow = q * 32 + lane;
        const int um = tm * kGemmBM + urow;
        const int n0 = nb + g * 8;
        const uint32_t off = part_u32 + static_cast<uint32_t>((g * kGemmBM + urow) * 32);
        float v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) v[jj] = 0.f;
#pragma unroll 1
        for (int s2 = 0; s2 < splits; s2 += 4) {                     // four peers in flight, summed in split order
          float t[4][8];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const bool okq = s2 + qq < splits;
            ld_cluster_f8(mapa_shared(off, static_cast<uint32_t>(okq ? s2 + qq : s2)), t[qq]);
            if (!okq) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) t[qq][jj] = 0.f;
            }
          }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] += t[qq][jj];
          }
        }
        if (!proceed || um >= PM || n0 >= PN) continue;
        float a0 = 0.f, a1 = 0.f;
        if (E.aux) {
          a0 = E.aux[static_cast<size_t>(um) * 2];
          a1 = E.aux[static_cast<size_t>(um) * 2 + 1];
        }
        ss += epi_group(E, X, v, um, n0, PN, a0, a1);
      }
      __syncwarp();
      cluster_arrive();
      cluster_wait();
    } else if (kVariant == 0) {
      // ---- split-K: publish this CTA's partial tile, wait for the other splits, finish a share of the tile ----
      // partial layout [split][group][row][8]: a warp's 32 rows of one group are 1 KB contiguous, so both the write
      // and the fix-up reads are fully coalesced
      float* const ws_tile = ws_base + static_cast<size_t>(tile) * splits * (kGemmBM * bn);
      float* mine = ws_tile + static_cast<size_t>(split) * (kGemmBM * bn) + static_cast<size_t>(row) * 8;
#pragma unroll 1
      for (int g = half; g < ngroups; g += 2) {
        uint32_t r[8];
        tmem_ld_32x8(taddr + g * 8, r);
        tmem_ld_wait();
        float pv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[j] = __uint_as_float(r[j]);
        stcg_f8(mine + g * (kGemmBM * 8), pv);
      }
      __threadfence();
      epi_named_barrier();
      int* const ctr = ctr_base + 2 * tile;
      if (epi_tid == 0) {
        atomicAdd(&ctr[0], 1);
        // all `splits` CTAs of this tile are resident (grid <= SM count, one CTA per SM): bounded spin
        long long t0 = 0;
        for (uint32_t spins = 1;; ++spins) {
          if (*reinterpret_cast<volatile int*>(&ctr[0]) >= splits) break;
          if ((spins & 255u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) {
              if (ctx.err) atomicExch(ctx.err, FND_DEV_TIMEOUT_SPLITK);
              break;
            }
          }
        }
        __threadfence();
      }
      epi_named_barrier();
      if (epi_tid == 0) FND_STAMP(5);
      // units: u = g * 4 + q  (q = 32-row quarter); unit u belongs to split (u % splits); this CTA's units are dealt
      // round-robin to its eight epilogue warps
      const int nunits = ngroups * 4;
      const size_t sstride = static_cast<size_t>(kGemmBM) * bn;
#pragma unroll 1
      for (int u = split + ew * splits; u < nunits; u += kGemmEpiWarps * splits) {
        const int g = u >> 2, q = u & 3;
        const int urow = q * 32 + lane;
        const int um = tm * kGemmBM + urow;
        const int n0 = nb + g * 8;
        float v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) v[jj] = 0.f;
        // fixed split order => deterministic sum; loads are issued eight splits at a time so the L2 round trips
        // overlap instead of serialising
        const float* src0 = ws_tile + static_cast<size_t>(g) * (kGemmBM * 8) + static_cast<size_t>(urow) * 8;
#pragma unroll 1
        for (int s2 = 0; s2 < splits; s2 += 8) {
          float t[8][8];
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const bool okq = s2 + qq < splits;
            ldcg_f8(src0 + static_cast<size_t>(okq ? s2 + qq : s2) * sstride, t[qq]);
            if (!okq) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) t[qq][jj] = 0.f;
            }
          }
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] += t[qq][jj];
          }
        }
        if (!proceed || um >= PM || n0 >= PN) continue;
        float a0 = 0.f, a1 = 0.f;
        if (E.aux) {
          a0 = E.aux[static_cast<size_t>(um) * 2];
          a1 = E.aux[static_cast<size_t>(um) * 2 + 1];
        }
        ss += epi_group(E, X, v, um, n0, PN, a0, a1);
      }
      // depart: the last CTA to finish reading re-arms both counters for the next launch (graph-replay safe)
      epi_named_barrier();
      if (epi_tid == 0) {
        const int old = atomicAdd(&ctr[1], 1);
        if (old == splits - 1) { ctr[0] = 0; ctr[1] = 0; }
      }
    }
    if (E.done_ctr) {
      __threadfence();
      epi_named_barrier();
      if (epi_tid == 0) atomicAdd(E.done_ctr, 1u);
    }
    if (E.sumsq_slots) {
      // every epilogue thread takes part => fixed order, deterministic
      ss = warp_sum(ss);
      if (lane == 0) red_smem[ew] = ss;
      epi_named_barrier();
      if (epi_tid == 0)
        E.sumsq_slots[local] = ((red_smem[0] + red_smem[1]) + (red_smem[2] + red_smem[3])) +
                               ((red_smem[4] + red_smem[5]) + (red_smem[6] + red_smem[7]));
    }
  }

  if (threadIdx.x == 64) FND_STAMP(6);
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) FND_STAMP(7);
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kGemmTmemCols);
  }
}

}  // namespace fnd
