// fnd_gemm_host.h — host-side construction of GemmProblem tables and the grouped launch.
#pragma once
#include "fnd_gemm.cuh"
#include "fnd_tmap.h"
#include <cstdlib>
#include <utility>
#include <vector>

namespace fnd {

// A GEMM operand as it lies in memory. Logical shape is [rows = M or N, contraction = K].
//   mn_major == false : memory is [rows][K]   (K contiguous)      e.g. activations / nn.Linear weights (fwd)
//   mn_major == true  : memory is [K][rows]   (rows contiguous)   e.g. W for dgrad, dY and X for wgrad
struct Operand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;   // may be null when ncombo == 1
  int pitch;                 // elements between consecutive memory rows
  bool mn_major;
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute (see fnd_common.cuh: every
// kernel of this library calls griddepcontrol.wait before touching data its predecessor may have produced).
// FND_NO_PDL=1 in the environment disables the attribute globally (debugging aid).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("FND_NO_PDL") ? 0 : 1;
  return v != 0;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(block), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
// The same launch with thread-block clusters of `cluster` consecutive CTAs (grid must be a multiple of it).
template <class... KArgs, class... Args>
inline cudaError_t launch_k_cluster(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl,
                                    int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(static_cast<unsigned>(block), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
// FND_CLUSTER_SPLITK=0 keeps the L2-workspace exchange (and the round-1 tile / split choices); fnd_debug_set_cluster_splitk
// flips the LAUNCH-time choice only, so that a test can run the same plan through both exchanges.
inline int& cluster_splitk_flag() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FND_CLUSTER_SPLITK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
inline bool cluster_splitk_enabled() { return cluster_splitk_flag() != 0; }

// Fills one problem; returns 0 or a negative error. `cta_begin` is assigned by the caller (finish_table).
inline int fill_problem(GemmProblem& p, const Operand& A, const Operand& B, int M, int N, int K, int bn, int splits,
                        int ncombo, unsigned long long hintA, unsigned long long hintB, float* ws, int* ctr,
                        const EpiParams& epi, int b_static = 0) {
  memset(&p, 0, sizeof(p));
  if (bn != 16 && bn != 32 && bn != 64 && bn != 128) return -10;
  if (B.mn_major && bn < 64) return -11;
  if (ncombo != 1 && ncombo != 3) return -12;
  if (ncombo == 3 && (!A.lo || !B.lo)) return -13;
  if (N % 8) return -15;                 // the epilogue works on 8-column groups
  // 256-bit epilogue accesses: fp32 operands whose pitch is a multiple of 8 floats must start 32-byte aligned
  if (epi.out_f32 && (epi.f32_pitch & 7) == 0 && (reinterpret_cast<uintptr_t>(epi.out_f32) & 31)) return -16;
  if (epi.out_pre && ((epi.pre_pitch & 7) || (reinterpret_cast<uintptr_t>(epi.out_pre) & 31))) return -16;
  if (epi.add_in && ((epi.add_pitch & 7) || (reinterpret_cast<uintptr_t>(epi.add_in) & 31))) return -16;
  if (epi.gate_z && ((epi.gate_pitch & 7) || (reinterpret_cast<uintptr_t>(epi.gate_z) & 31))) return -16;
  p.M = M; p.N = N; p.K = K;
  p.bn = bn;
  p.ncombo = ncombo;
  p.b_static = b_static;
  p.stage_bytes = kGemmStageBytesA + bn * kGemmBK * 2;
  p.nstages = kGemmOperandBudget / p.stage_bytes;
  if (p.nstages > kGemmMaxStages) p.nstages = kGemmMaxStages;
  p.tiles_m = ceil_div(M, kGemmBM);
  p.tiles_n = ceil_div(N, bn);
  p.kb_total = ceil_div(K, kGemmBK);
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = ceil_div(p.kb_total, splits);
  p.splits = ceil_div(p.kb_total, p.kb_per_split);   // every split non-empty
  if (p.splits > 1 && (!ws || !ctr)) return -14;
  p.cta_count = p.tiles_m * p.tiles_n * p.splits;
  if (p.nstages > p.kb_per_split * ncombo) p.nstages = p.kb_per_split * ncombo;   // never deeper than the k-loop
  p.hintA = hintA; p.hintB = hintB;
  p.splitk_ws = ws; p.splitk_ctr = ctr;
  p.epi = epi;
  p.epi.plain_f32 = (epi.out_f32 && !epi.bias && !epi.aux && !epi.add_in && !epi.out_pre && !epi.act && !epi.gate_z &&
                     !epi.out_hi && epi.drop_p == 0.f && (epi.f32_pitch & 7) == 0) ? 1 : 0;
  if (p.epi.plain_f32 && p.nstages * p.stage_bytes >= kGemmEpiWarps * 32 * 36 * 4) p.epi.plain_f32 = 2;   // room to stage the tile
  if (!p.epi.plain_f32 || bn < 64) p.epi.out_bf = nullptr;    // the bf16 mirror is written by the store-only epilogues only
  for (int h = 0; h < (ncombo == 3 ? 2 : 1); ++h) {
    const void* a = h ? static_cast<const void*>(A.lo) : static_cast<const void*>(A.hi);
    const void* b = h ? static_cast<const void*>(B.lo) : static_cast<const void*>(B.hi);
    int r;
    if (!A.mn_major) r = encode_bf16_2d(&p.tmA[h], a, K, M, A.pitch, 64, kGemmBM);
    else             r = encode_bf16_2d(&p.tmA[h], a, M, K, A.pitch, 64, kGemmBK);
    if (r) return r;
    if (!B.mn_major) r = encode_bf16_2d(&p.tmB[h], b, K, N, B.pitch, 64, bn);
    else             r = encode_bf16_2d(&p.tmB[h], b, N, K, B.pitch, 64, kGemmBK);
    if (r) return r;
  }
  if (ncombo == 1) { p.tmA[1] = p.tmA[0]; p.tmB[1] = p.tmB[0]; }
  return 0;
}

inline size_t splitk_ws_floats(int M, int N, int bn, int splits) {
  return static_cast<size_t>(ceil_div(M, kGemmBM)) * ceil_div(N, bn) * splits * kGemmBM * bn;
}

// Assigns cta_begin prefix sums; returns the grid size.
inline int finish_table(GemmProblem* probs, int n) {
  int c = 0;
  for (int i = 0; i < n; ++i) { probs[i].cta_begin = c; c += probs[i].cta_count; }
  return c;
}

// Opt the GEMM kernel into its dynamic shared-memory size (per device; called at bind time so that no attribute
// call ever happens inside a stream capture).
inline cudaError_t init_gemm_attrs() {
  cudaError_t e = cudaFuncSetAttribute(fnd_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fnd_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
}

// Dynamic shared memory one launch of `host_table` needs: header + the deepest ring among its problems (+ align slack).
inline int gemm_smem_bytes(const GemmProblem* host_table, int nprob) {
  int ring = 0;
  for (int i = 0; i < nprob; ++i) {
    const int r = host_table[i].nstages * host_table[i].stage_bytes;
    ring = r > ring ? r : ring;
  }
  return ring + 1024 + kGemmSmemHeader;
}

// kind: 0 = forward (A K-major, B K-major); 1 = dgrad (A K-major, B MN-major); 2 = wgrad (both MN-major);
// 3 = (A MN-major, B K-major). `host_table` is copied into the kernel's parameter space at launch.
// The light variant (two CTAs per SM, trailing finalize CTAs) serves wgrad launches without split-K whose ring is
// shallow enough for two CTAs to share an SM (contraction = batch <= 192).
inline bool gemm_launch_is_light(int kind, const GemmProblem* host_table, int nprob) {
  bool light = kind == 2 && gemm_smem_bytes(host_table, nprob) <= 100 * 1024;
  for (int i = 0; i < nprob; ++i) light = light && host_table[i].splits == 1;
  return light;
}

// `fin` (optional): finalize jobs run by `fin_ctas` trailing CTAs of the launch (light variant, see fnd_gemm.cuh).
// The light variant is chosen for launches without split-K whose ring is shallow enough for two CTAs per SM.
inline cudaError_t launch_gemm(int kind, const GemmProblem* host_table, int nprob, int grid, RunCtx ctx,
                               cudaStream_t st, bool pdl = false, const FinParams* fin = nullptr, int fin_ctas = 0) {
  if (grid <= 0) return cudaSuccess;
  if (nprob < 1 || nprob > kGemmTableCap || kind < 0 || kind > 3) return cudaErrorInvalidValue;
  GemmTableP t;
  memcpy(t.p, host_table, sizeof(GemmProblem) * nprob);
  t.nprob = nprob;
  t.a_mn = (kind == 2 || kind == 3) ? 1 : 0;
  t.b_mn = (kind == 1 || kind == 2) ? 1 : 0;
  t.gemm_ctas = grid;
  FinParams f;
  memset(&f, 0, sizeof(f));
  if (fin) f = *fin;
  const int smem = gemm_smem_bytes(host_table, nprob);
  if (!gemm_launch_is_light(kind, host_table, nprob) && fin_ctas > 0) return cudaErrorInvalidValue;
  const bool light = gemm_launch_is_light(kind, host_table, nprob);
  t.cluster_k = 0;
  if (light) return launch_k(fnd_gemm_kernel<1>, grid + fin_ctas, kGemmThreads, smem, st, pdl, t, ctx, f);
  // (16-CTA clusters for the 16 k-splits of gemm_fuse0 — non-portable size — were measured too: no change in the step.)
  // Split-K launches whose problems all use the same 2 / 4 / 8 splits run as clusters of the k-splits of a tile and exchange
  // their partial tiles through distributed shared memory (every problem's CTA count is a multiple of its splits, so the
  // clusters line up with the tiles); a partial tile must fit in the problem's (dead) operand ring.
  const int S = host_table[0].splits;
  bool ck = cluster_splitk_enabled() && (S == 2 || S == 4 || S == 8) && grid % S == 0;
  for (int i = 0; i < nprob && ck; ++i)
    ck = host_table[i].splits == S && host_table[i].cta_begin % S == 0 &&
         host_table[i].nstages * host_table[i].stage_bytes >= kGemmBM * host_table[i].bn * 4;
  if (ck) {
    t.cluster_k = 1;
    return launch_k_cluster(fnd_gemm_kernel<0>, grid, kGemmThreads, smem, st, pdl, S, t, ctx, f);
  }
  return launch_k(fnd_gemm_kernel<0>, grid, kGemmThreads, smem, st, pdl, t, ctx, f);
}

}  // namespace fnd
