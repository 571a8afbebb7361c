// fnd_seq_api.cu — C-ABI of the sequence front-end (see include/fnd_seq_b200.h for the contract of every entry point).
#include "../../include/fnd_seq_b200.h"
#include "fnd_seq_attn.cuh"
#include "fnd_seq_attn_bwd.cuh"
#include "fnd_seq_bwd_rows.cuh"
#include "fnd_seq_gemm.cuh"
#include "fnd_seq_gemm2.cuh"
#include "fnd_seq_rows.cuh"
#include "fnd_tmap.h"
#include <math.h>
#include <stdlib.h>

using namespace fnd;

#define SEQ_CUDA_OK(expr)                                        \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return -1000 - static_cast<int>(_e);  \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int seq_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Clusters of the CTA-pair GEMM that can be resident at once (one CTA per SM: normally 74); 0 = pair kernel unavailable.
static int g_pair_clusters = 0;

static long long* g_gemm_stamps = nullptr;      // probe aid (fnd_seq_debug_gemm_stamps)
static long long* g_attn_stamps = nullptr;      // probe aid (fnd_seq_debug_attn_stamps)

extern "C" {

int fnd_seq_debug_attn_stamps(long long* stamps) {
  g_attn_stamps = stamps;
  return 0;
}

int fnd_seq_debug_gemm_stamps(long long* stamps) {
  g_gemm_stamps = stamps;
  return 0;
}

int fnd_seq_init(void) {
  static bool done = false;
  if (done) return 0;
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqGemmSmemMax));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_gemm2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqGemmSmemMax));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_gemm2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqGemmSmemMax));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_gemm2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqGemmSmemMax));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_gemm2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqGemmSmemMax));
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(seq_num_sms() & ~1), 1, 1);
    cfg.blockDim = dim3(kSeqGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = kSeqGemmSmemMax;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, seq_gemm2_kernel<1, false>, &cfg) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = seq_num_sms() / 2;                  // one CTA per SM by shared memory: every SM pair holds one cluster
    }
    const char* e = getenv("FND_SEQ_GEMM_PAIR");
    g_pair_clusters = (e && atoi(e) == 0) ? 0 : nc;
  }
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_layernorm_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 256 * 4));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 512 * 4));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
  SEQ_CUDA_OK(cudaFuncSetAttribute(seq_layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4));
  done = true;
  return 0;
}

int fnd_seq_pair_clusters(void) {
  { int r = fnd_seq_init(); if (r) return 0; }
  return g_pair_clusters;
}

int fnd_seq_cast_bf16(const float* x, void* y_bf16, long long n, void* stream) {
  if (!x || !y_bf16 || n < 0 || (n & 7) || !aligned16(x) || !aligned16(y_bf16)) return -1;
  if (n == 0) return 0;
  const size_t n8 = static_cast<size_t>(n) >> 3;
  const int grid = static_cast<int>(n8 / 256 + 1 < static_cast<size_t>(seq_num_sms()) * 8 ? n8 / 256 + 1 : static_cast<size_t>(seq_num_sms()) * 8);
  seq_cast_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y_bf16), n8);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_linear(const void* a_bf16, int a_pitch, const void* w_bf16, int w_pitch, const float* bias,
                   const void* resid_bf16, int resid_pitch, int act, void* out_bf16, int out_pitch, float* out_f32,
                   int f32_pitch, int M, int N, int K, int* err_flag, void* stream) {
  if (!a_bf16 || !w_bf16 || (!out_bf16 && !out_f32) || M <= 0 || N <= 0 || K <= 0) return -1;
  if ((N & 7) || (K & 7) || (a_pitch & 7) || (w_pitch & 7) || a_pitch < K || w_pitch < K) return -2;
  if (out_bf16 && ((out_pitch & 7) || out_pitch < N || !aligned16(out_bf16))) return -3;
  if (out_f32 && ((f32_pitch & 3) || f32_pitch < N || !aligned16(out_f32))) return -3;
  if (resid_bf16 && ((resid_pitch & 7) || resid_pitch < N || !aligned16(resid_bf16))) return -3;
  if (bias && !aligned16(bias)) return -3;
  { int r = fnd_seq_init(); if (r) return r; }
  SeqGemmParams P;
  memset(&P, 0, sizeof(P));
  P.M = M; P.N = N; P.K = K;
  P.splits = 1; P.kb_per_split = cdiv(K, kSeqGemmBK);
  // Large bf16-output problems: 256 x 256 tiles on CTA pairs (fnd_seq_gemm2.cuh) when there is at least one tile per pair
  if (g_pair_clusters > 0 && out_bf16 && !out_f32 && N >= kSeqGemm2BN && (N & 63) == 0 &&
      static_cast<long long>(cdiv(M, 2 * kSeqGemmBM)) * cdiv(N, kSeqGemm2BN) >= g_pair_clusters) {
    P.bn = kSeqGemm2BN;
    P.tiles_m = cdiv(M, 2 * kSeqGemmBM);
    P.tiles_n = cdiv(N, kSeqGemm2BN);
    // one 64-deep k-box per stage (six 32 KB stages, five beside the residual slabs); FND_SEQ_GEMM_KB=2 selects 128-deep
    // stages for the residual-free GEMMs — measured no faster (fnd_seq_gemm2.cuh)
    static const char* ekb = getenv("FND_SEQ_GEMM_KB");
    const int kkb = (ekb && atoi(ekb) == 2 && !resid_bf16) ? 2 : 1;
    P.kblocks = cdiv(K, kSeqGemmBK * kkb);
    P.stage_bytes = kkb * kSeqGemm2StageBytes;
    const int rbuf_bytes = resid_bf16 ? 2 * kSeqGemmStageOutBytes : 0;
    P.nstages = (kSeqGemmRingBudget - rbuf_bytes) / P.stage_bytes;
    if (P.nstages > kSeqGemmMaxStages) P.nstages = kSeqGemmMaxStages;
    P.bias = bias;
    P.resid = static_cast<const __nv_bfloat16*>(resid_bf16); P.resid_pitch = resid_pitch;
    P.act = act;
    P.out_bf = static_cast<__nv_bfloat16*>(out_bf16); P.out_pitch = out_pitch;
    P.err = err_flag;
    { static const char* e = getenv("FND_SEQ_DBG_MMAS"); P.dbg_mmas = e ? atoi(e) : 0; }
    P.dbg = g_gemm_stamps;
    int r = encode_bf16_2d(&P.tmA, a_bf16, K, M, a_pitch, 64, kSeqGemmBM);
    if (r) return r;
    r = encode_bf16_2d(&P.tmB, w_bf16, K, N, w_pitch, 64, kSeqGemm2BN / 2);
    if (r) return r;
    r = encode_bf16_2d(&P.tmC, out_bf16, N, M, out_pitch, 64, kSeqGemmBM);
    if (r) return r;
    if (resid_bf16) {
      r = encode_bf16_2d(&P.tmR, resid_bf16, N, M, resid_pitch, 64, kSeqGemmBM);
      if (r) return r;
    }
    const int ntiles = P.tiles_m * P.tiles_n;
    const int nclusters = ntiles < g_pair_clusters ? ntiles : g_pair_clusters;
    const size_t smem = static_cast<size_t>(P.nstages) * P.stage_bytes + 2 * kSeqGemmStageOutBytes + rbuf_bytes + kSeqGemmHeader + 1024;
    const cudaStream_t st2 = reinterpret_cast<cudaStream_t>(stream);
    const bool dbg = P.dbg != nullptr || P.dbg_mmas != 0;
    if (kkb == 2) {
      if (dbg) seq_gemm2_kernel<2, true><<<2 * nclusters, kSeqGemmThreads, smem, st2>>>(P);
      else seq_gemm2_kernel<2, false><<<2 * nclusters, kSeqGemmThreads, smem, st2>>>(P);
    } else {
      if (dbg) seq_gemm2_kernel<1, true><<<2 * nclusters, kSeqGemmThreads, smem, st2>>>(P);
      else seq_gemm2_kernel<1, false><<<2 * nclusters, kSeqGemmThreads, smem, st2>>>(P);
    }
    SEQ_CUDA_OK(cudaGetLastError());
    return 0;
  }
  P.tiles_m = cdiv(M, kSeqGemmBM);
  // widest tile that still gives every SM work; 256 columns keep one UMMA busy for 128 cycles
  int bn = 256;
  const int sms = seq_num_sms();
  if (N % 256 != 0 && N <= 128) bn = N <= 64 ? 64 : 128;
  while (bn > 64 && P.tiles_m * cdiv(N, bn) < sms) bn >>= 1;
  P.bn = bn;
  P.tiles_n = cdiv(N, bn);
  P.kblocks = cdiv(K, kSeqGemmBK);
  P.stage_bytes = kSeqGemmBM * kSeqGemmBK * 2 + bn * kSeqGemmBK * 2;
  const int rbuf_bytes = (resid_bf16 && out_bf16) ? 2 * kSeqGemmStageOutBytes : 0;      // residual slabs, prefetched two ahead
  P.nstages = (kSeqGemmRingBudget - rbuf_bytes) / P.stage_bytes;
  if (P.nstages > kSeqGemmMaxStages) P.nstages = kSeqGemmMaxStages;
  P.bias = bias;
  P.resid = static_cast<const __nv_bfloat16*>(resid_bf16); P.resid_pitch = resid_pitch;
  P.act = act;
  P.out_bf = static_cast<__nv_bfloat16*>(out_bf16); P.out_pitch = out_pitch;
  P.out_f32 = out_f32; P.f32_pitch = f32_pitch;
  P.err = err_flag;
  int r = encode_bf16_2d(&P.tmA, a_bf16, K, M, a_pitch, 64, kSeqGemmBM);
  if (r) return r;
  r = encode_bf16_2d(&P.tmB, w_bf16, K, N, w_pitch, 64, bn);
  if (r) return r;
  if (out_bf16) {
    r = encode_bf16_2d(&P.tmC, out_bf16, N, M, out_pitch, 64, kSeqGemmBM);
    if (r) return r;
  }
  if (resid_bf16) {
    r = encode_bf16_2d(&P.tmR, resid_bf16, N, M, resid_pitch, 64, kSeqGemmBM);
    if (r) return r;
  }
  const int ntiles = P.tiles_m * P.tiles_n;
  const int grid = ntiles < sms ? ntiles : sms;
  const size_t smem = static_cast<size_t>(P.nstages) * P.stage_bytes + 2 * kSeqGemmStageOutBytes + rbuf_bytes + kSeqGemmHeader + 1024;
  seq_gemm_kernel<<<grid, kSeqGemmThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_layernorm(const void* x_bf16, int x_pitch, const void* resid_bf16, int resid_pitch, const float* gamma,
                      const float* beta, float eps, void* y_bf16, int y_pitch, int M, int d, void* stream) {
  if (!x_bf16 || !gamma || !beta || !y_bf16 || M <= 0 || d <= 0) return -1;
  if ((d & 7) || d > kLnMaxChunks * 256 || (x_pitch & 7) || (y_pitch & 7) || x_pitch < d || y_pitch < d) return -2;
  if (resid_bf16 && ((resid_pitch & 7) || resid_pitch < d || !aligned16(resid_bf16))) return -2;
  if (!aligned16(x_bf16) || !aligned16(y_bf16) || !aligned16(gamma) || !aligned16(beta)) return -3;
  LnParams P{static_cast<const __nv_bfloat16*>(x_bf16), x_pitch, static_cast<const __nv_bfloat16*>(resid_bf16), resid_pitch,
             gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16), y_pitch, M, d};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // resident warps walk the row list (gamma / beta stay in registers): a few waves of CTAs, never more than the rows need
  const int want = cdiv(M, 8);
  const int cap = seq_num_sms() * (d <= 1024 ? 4 : 2);
  const int grid = want < cap ? want : cap;
  if (d <= 256) seq_layernorm_kernel<1><<<grid, 256, 0, st>>>(P);
  else if (d <= 512) seq_layernorm_kernel<2><<<grid, 256, 0, st>>>(P);
  else if (d <= 1024) seq_layernorm_kernel<4><<<grid, 256, 0, st>>>(P);
  else seq_layernorm_kernel<8><<<grid, 256, 0, st>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_coattn_forward(const void* q_bf16, int q_pitch, int q_col0, const void* k_bf16, int k_pitch, int k_col0,
                           const void* v_bf16, int v_pitch, int v_col0, const int* kv_len, const unsigned char* kv_mask,
                           int B, int H, int Lq, int Lk, float scale, void* out_bf16, int out_pitch, float* lse,
                           int* err_flag, void* stream) {
  if (!q_bf16 || !k_bf16 || !v_bf16 || !out_bf16 || B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return -1;
  if ((q_pitch & 7) || (k_pitch & 7) || (v_pitch & 7) || (out_pitch & 7) || (q_col0 & 7) || (k_col0 & 7) || (v_col0 & 7)) return -2;
  if (q_col0 + H * kAttnD > q_pitch || k_col0 + H * kAttnD > k_pitch || v_col0 + H * kAttnD > v_pitch || H * kAttnD > out_pitch) return -2;
  if (!aligned16(out_bf16) || !(scale > 0.f) || H > 65535 || B > 65535) return -3;
  { int r = fnd_seq_init(); if (r) return r; }
  AttnParams P;
  memset(&P, 0, sizeof(P));
  int r = encode_bf16_3d(&P.tmQ, q_bf16, static_cast<uint64_t>(q_pitch), Lq, B, q_pitch, kAttnBQ);
  if (r) return r;
  r = encode_bf16_3d(&P.tmK, k_bf16, static_cast<uint64_t>(k_pitch), Lk, B, k_pitch, kAttnBK);
  if (r) return r;
  r = encode_bf16_3d(&P.tmV, v_bf16, static_cast<uint64_t>(v_pitch), Lk, B, v_pitch, kAttnBK);
  if (r) return r;
  r = encode_bf16_3d(&P.tmO, out_bf16, static_cast<uint64_t>(H) * kAttnD, Lq, B, out_pitch, kAttnBQ);
  if (r) return r;
  P.B = B; P.H = H; P.Lq = Lq; P.Lk = Lk;
  P.q_col0 = q_col0; P.k_col0 = k_col0; P.v_col0 = v_col0;
  P.kv_len = kv_len; P.kv_mask = kv_mask;
  P.scale = scale;
  P.scale_log2 = scale * 1.44269504088896340736f;
  P.out = static_cast<__nv_bfloat16*>(out_bf16); P.out_pitch = out_pitch;
  P.lse = lse;
  P.err = err_flag;
  // persistent: one resident CTA per SM walks the (sample, head, query-tile) work list
  const long long nwork = static_cast<long long>(cdiv(Lq, kAttnItemQ)) * H * B;
  if (nwork > 0x7fffffffLL) return -3;
  const int grid = nwork < 1LL * seq_num_sms() ? static_cast<int>(nwork) : seq_num_sms();
  // experiment knobs (tools/attn_probe.py sweeps them); the defaults are the measured best
  int poly = kAttnDefaultPoly, skew = kAttnDefaultSkewNs;
  if (const char* e = getenv("FND_ATTN_POLY")) { poly = atoi(e); if (poly < 0 || poly > 3) poly = kAttnDefaultPoly; }
  if (const char* e = getenv("FND_ATTN_SKEW_NS")) { skew = atoi(e); if (skew < 0) skew = 0; }
  P.skew_ns = skew;
  P.pingpong = kAttnDefaultPingPong;
  if (const char* e = getenv("FND_ATTN_PINGPONG")) P.pingpong = atoi(e) != 0;
  if (const char* e = getenv("FND_ATTN_DBG_NOEXP")) P.dbg_noexp = atoi(e) != 0;
  P.dbg = g_attn_stamps;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (poly) {
    case 0: seq_attn_fwd_kernel<0><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(P); break;
    case 1: seq_attn_fwd_kernel<1><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(P); break;
    case 2: seq_attn_fwd_kernel<2><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(P); break;
    default: seq_attn_fwd_kernel<3><<<grid, kAttnThreads, kAttnSmemBytes, st>>>(P); break;
  }
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_masked_mean_pool(const void* x_bf16, int x_pitch, const unsigned char* mask, const int* len, int B, int L,
                             int d, float* out_f32, int f32_pitch, void* out_bf16, int bf_pitch, void* stream) {
  if (!x_bf16 || (!out_f32 && !out_bf16) || B <= 0 || L <= 0 || d <= 0) return -1;
  if ((d & 7) || (x_pitch & 7) || x_pitch < d || !aligned16(x_bf16) || B > 65535) return -2;
  PoolParams P{static_cast<const __nv_bfloat16*>(x_bf16), x_pitch, mask, len, B, L, d, out_f32, f32_pitch,
               static_cast<__nv_bfloat16*>(out_bf16), bf_pitch};
  dim3 grid(static_cast<unsigned>(cdiv(d, 64)), static_cast<unsigned>(B));
  seq_masked_mean_pool_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------- backward pass

// dW[N_out, K_in] = dY[tokens, N_out]^T X[tokens, K_in]: split plan shared by the workspace query and the launch
static void wgrad_plan(int tokens, int n_out, int k_in, int* bn, int* splits, int* kb_per_split) {
  *bn = (k_in % 256 == 0 || k_in > 256) ? 256 : (k_in > 64 ? 128 : 64);
  const int tiles = cdiv(n_out, kSeqGemmBM) * cdiv(k_in, *bn);
  const int kblocks = cdiv(tokens, kSeqGemmBK);
  int s = cdiv(4 * seq_num_sms(), tiles);                 // >= 4 waves of items
  const int smax = kblocks / 16 > 0 ? kblocks / 16 : 1;   // at least 16 k-blocks (1024 tokens) per item
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  *kb_per_split = cdiv(kblocks, s);
  *splits = cdiv(kblocks, *kb_per_split);                 // every split non-empty
}

size_t fnd_seq_wgrad_workspace(int tokens, int n_out, int k_in) {
  if (tokens <= 0 || n_out <= 0 || k_in <= 0) return 0;
  int bn, splits, kbps;
  wgrad_plan(tokens, n_out, k_in, &bn, &splits, &kbps);
  return splits > 1 ? static_cast<size_t>(splits) * n_out * k_in * sizeof(float) : 16;
}

int fnd_seq_wgrad(const void* dy_bf16, int dy_pitch, const void* x_bf16, int x_pitch, int tokens, int n_out, int k_in,
                  float* dw, int dw_pitch, void* workspace, size_t workspace_bytes, int* err_flag, void* stream) {
  if (!dy_bf16 || !x_bf16 || !dw || !workspace || tokens <= 0 || n_out <= 0 || k_in <= 0) return -1;
  if ((n_out & 7) || (k_in & 7) || (dy_pitch & 7) || (x_pitch & 7) || dy_pitch < n_out || x_pitch < k_in) return -2;
  if ((dw_pitch & 3) || dw_pitch < k_in || !aligned16(dw) || !aligned16(workspace)) return -3;
  if (workspace_bytes < fnd_seq_wgrad_workspace(tokens, n_out, k_in)) return -4;
  { int r = fnd_seq_init(); if (r) return r; }
  SeqGemmParams P;
  memset(&P, 0, sizeof(P));
  int bn;
  wgrad_plan(tokens, n_out, k_in, &bn, &P.splits, &P.kb_per_split);
  P.mn = 1;
  P.M = n_out; P.N = k_in; P.K = tokens;
  P.bn = bn;
  P.tiles_m = cdiv(n_out, kSeqGemmBM);
  P.tiles_n = cdiv(k_in, bn);
  P.kblocks = cdiv(tokens, kSeqGemmBK);
  P.stage_bytes = kSeqGemmBM * kSeqGemmBK * 2 + bn * kSeqGemmBK * 2;
  P.nstages = kSeqGemmRingBudget / P.stage_bytes;
  if (P.nstages > kSeqGemmMaxStages) P.nstages = kSeqGemmMaxStages;
  P.err = err_flag;
  const bool split = P.splits > 1;
  // partial tiles are dense [n_out, k_in]; without a split the kernel writes dw directly
  P.out_f32 = split ? static_cast<float*>(workspace) : dw;
  P.f32_pitch = split ? k_in : dw_pitch;
  P.split_stride = static_cast<long long>(n_out) * k_in;
  if (!split && (dw_pitch & 3)) return -3;
  int r = encode_bf16_2d(&P.tmA, dy_bf16, n_out, tokens, dy_pitch, 64, kSeqGemmBK);     // dY as [tokens][n_out]: MN-major A
  if (r) return r;
  r = encode_bf16_2d(&P.tmB, x_bf16, k_in, tokens, x_pitch, 64, kSeqGemmBK);            // X as [tokens][k_in]: MN-major B
  if (r) return r;
  const int sms = seq_num_sms();
  const long long nitems = static_cast<long long>(P.tiles_m) * P.tiles_n * P.splits;
  const int grid = nitems < sms ? static_cast<int>(nitems) : sms;
  const size_t smem = static_cast<size_t>(P.nstages) * P.stage_bytes + 2 * kSeqGemmStageOutBytes + kSeqGemmHeader + 1024;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  seq_gemm_kernel<<<grid, kSeqGemmThreads, smem, st>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  if (split) {
    if (dw_pitch != k_in) return -3;                     // the fixed-order reduction writes a dense matrix
    const long long n = static_cast<long long>(n_out) * k_in;
    seq_reduce_partials_kernel<<<static_cast<unsigned>((n + 127) / 128), 256, 0, st>>>(static_cast<const float*>(workspace), P.splits,
                                                                                          static_cast<int>(n), dw);
    SEQ_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

static int ln_bwd_grid(int M, int d) {
  const int want = cdiv(M, 8);
  const int cap = seq_num_sms() * (d <= 512 ? 2 : 1);
  return want < cap ? want : cap;
}

size_t fnd_seq_layernorm_backward_workspace(int M, int d) {
  if (M <= 0 || d <= 0) return 0;
  return static_cast<size_t>(ln_bwd_grid(M, d)) * 2 * d * sizeof(float);
}

int fnd_seq_layernorm_backward(const void* t_bf16, int t_pitch, const void* dy_bf16, int dy_pitch, const float* gamma,
                               float eps, void* dt_bf16, int dt_pitch, float* dgamma, float* dbeta, int M, int d,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!t_bf16 || !dy_bf16 || !gamma || !dt_bf16 || !dgamma || !dbeta || !workspace || M <= 0 || d <= 0) return -1;
  if ((d & 7) || d > kLnMaxChunks * 256 || (t_pitch & 7) || (dy_pitch & 7) || (dt_pitch & 7) || t_pitch < d || dy_pitch < d || dt_pitch < d) return -2;
  if (!aligned16(t_bf16) || !aligned16(dy_bf16) || !aligned16(dt_bf16) || !aligned16(gamma) || !aligned16(workspace) || !aligned16(dgamma) || !aligned16(dbeta)) return -3;
  if (dbeta != dgamma + d) return -3;             // the two gradients are reduced as ONE 2d-wide row: pass adjacent buffers
  if (workspace_bytes < fnd_seq_layernorm_backward_workspace(M, d)) return -4;
  { int r = fnd_seq_init(); if (r) return r; }
  LnBwdParams P{static_cast<const __nv_bfloat16*>(t_bf16), t_pitch, static_cast<const __nv_bfloat16*>(dy_bf16), dy_pitch, gamma, eps,
                static_cast<__nv_bfloat16*>(dt_bf16), dt_pitch, static_cast<float*>(workspace), M, d};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_bwd_grid(M, d);
  const size_t smem = static_cast<size_t>(8) * 2 * d * sizeof(float);
  if (d <= 256) seq_layernorm_bwd_kernel<1><<<grid, 256, smem, st>>>(P);
  else if (d <= 512) seq_layernorm_bwd_kernel<2><<<grid, 256, smem, st>>>(P);
  else if (d <= 1024) seq_layernorm_bwd_kernel<4><<<grid, 256, smem, st>>>(P);
  else seq_layernorm_bwd_kernel<8><<<grid, 256, smem, st>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  seq_reduce_partials_kernel<<<cdiv(2 * d, 128), 256, 0, st>>>(static_cast<const float*>(workspace), grid, 2 * d, dgamma);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

static int colsum_slices(int M) {
  const int want = cdiv(M, 64);
  const int cap = seq_num_sms() * 2;
  return want < cap ? want : cap;
}

size_t fnd_seq_colsum_workspace(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return static_cast<size_t>(colsum_slices(M)) * N * sizeof(float);
}

int fnd_seq_colsum(const void* x_bf16, int x_pitch, int M, int N, float* out, void* workspace, size_t workspace_bytes,
                   void* stream) {
  if (!x_bf16 || !out || !workspace || M <= 0 || N <= 0) return -1;
  if ((N & 7) || (x_pitch & 7) || x_pitch < N || !aligned16(x_bf16) || !aligned16(out) || !aligned16(workspace)) return -2;
  if (workspace_bytes < fnd_seq_colsum_workspace(M, N)) return -4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int slices = colsum_slices(M);
  ColsumParams P{static_cast<const __nv_bfloat16*>(x_bf16), x_pitch, M, N, static_cast<float*>(workspace)};
  dim3 grid(static_cast<unsigned>(cdiv(N, 256)), static_cast<unsigned>(slices));
  seq_colsum_kernel<<<grid, 256, 0, st>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  seq_reduce_partials_kernel<<<cdiv(N, 128), 256, 0, st>>>(static_cast<const float*>(workspace), slices, N, out);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_masked_mean_pool_backward(const float* dpooled, int dp_pitch, const unsigned char* mask, const int* len, int B,
                                      int L, int d, void* dx_bf16, int dx_pitch, void* stream) {
  if (!dpooled || !dx_bf16 || B <= 0 || L <= 0 || d <= 0) return -1;
  if ((d & 7) || (dx_pitch & 7) || dx_pitch < d || (dp_pitch & 3) || dp_pitch < d || !aligned16(dx_bf16) || !aligned16(dpooled)) return -2;
  PoolBwdParams P{dpooled, dp_pitch, mask, len, B, L, d, static_cast<__nv_bfloat16*>(dx_bf16), dx_pitch};
  const long long M = static_cast<long long>(B) * L;
  const long long want = (M + 7) / 8;
  const int cap = seq_num_sms() * 8;
  const int grid = want < cap ? static_cast<int>(want) : cap;
  seq_masked_mean_pool_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t fnd_seq_coattn_backward_workspace(int B, int H, int Lq) {
  if (B <= 0 || H <= 0 || Lq <= 0) return 0;
  const size_t Lqp = static_cast<size_t>(cdiv(Lq, 64)) * 64;
  return 2 * static_cast<size_t>(B) * H * Lqp * sizeof(float);
}

int fnd_seq_coattn_backward(const void* q_bf16, int q_pitch, int q_col0, const void* k_bf16, int k_pitch, int k_col0,
                            const void* v_bf16, int v_pitch, int v_col0, const void* o_bf16, int o_pitch,
                            const void* do_bf16, int do_pitch, const float* lse, const int* kv_len,
                            const unsigned char* kv_mask, int B, int H, int Lq, int Lk, float scale, void* dq_bf16,
                            int dq_pitch, int dq_col0, void* dk_bf16, int dk_pitch, int dk_col0, void* dv_bf16,
                            int dv_pitch, int dv_col0, void* workspace, size_t workspace_bytes, int* err_flag,
                            void* stream) {
  if (!q_bf16 || !k_bf16 || !v_bf16 || !o_bf16 || !do_bf16 || !lse || !dq_bf16 || !dk_bf16 || !dv_bf16 || !workspace) return -1;
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || H > 65535 || B > 65535 || !(scale > 0.f)) return -1;
  const int pitches[8] = {q_pitch, k_pitch, v_pitch, o_pitch, do_pitch, dq_pitch, dk_pitch, dv_pitch};
  for (int i = 0; i < 8; ++i) if (pitches[i] & 7) return -2;
  const int col0s[6] = {q_col0, k_col0, v_col0, dq_col0, dk_col0, dv_col0};
  for (int i = 0; i < 6; ++i) if (col0s[i] & 7) return -2;
  const int hd = H * kAttnD;
  if (q_col0 + hd > q_pitch || k_col0 + hd > k_pitch || v_col0 + hd > v_pitch || hd > o_pitch || hd > do_pitch ||
      dq_col0 + hd > dq_pitch || dk_col0 + hd > dk_pitch || dv_col0 + hd > dv_pitch) return -2;
  if (!aligned16(dq_bf16) || !aligned16(dk_bf16) || !aligned16(dv_bf16) || !aligned16(o_bf16) || !aligned16(do_bf16) || !aligned16(workspace)) return -3;
  if (workspace_bytes < fnd_seq_coattn_backward_workspace(B, H, Lq)) return -4;
  { int r = fnd_seq_init(); if (r) return r; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Lqp = cdiv(Lq, 64) * 64;
  float* Dp = static_cast<float*>(workspace);
  float* lse2p = Dp + static_cast<size_t>(B) * H * Lqp;
  {
    AttnBwdPrepParams R{static_cast<const __nv_bfloat16*>(o_bf16), o_pitch, static_cast<const __nv_bfloat16*>(do_bf16), do_pitch, lse, B, H, Lq, Lqp, Dp, lse2p};
    const long long rows = static_cast<long long>(B) * Lqp;
    const long long want = (rows + 7) / 8;
    const int cap = seq_num_sms() * 8;
    seq_attn_bwd_prep_kernel<<<want < cap ? static_cast<int>(want) : cap, 256, 0, st>>>(R);
    SEQ_CUDA_OK(cudaGetLastError());
  }
  const int sms = seq_num_sms();
  AttnBwdParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.H = H; P.Lq = Lq; P.Lk = Lk; P.Lqp = Lqp;
  P.kv_len = kv_len; P.kv_mask = kv_mask;
  P.scale = scale; P.scale_log2 = scale * 1.44269504088896340736f;
  P.lse2p = lse2p; P.Dp = Dp;
  P.err = err_flag;
  P.pingpong = kAttnBwdDefaultPingPong;
  if (const char* e = getenv("FND_ATTN_BWD_PINGPONG")) P.pingpong = atoi(e) != 0;
  int r;
  // ---- dQ: resident Q / dO tiles (128 rows), streamed K / V blocks (64 rows) ----
  if ((r = encode_bf16_3d(&P.tmR0, q_bf16, static_cast<uint64_t>(q_pitch), Lq, B, q_pitch, kAttnBQ))) return r;
  if ((r = encode_bf16_3d(&P.tmR1, do_bf16, static_cast<uint64_t>(do_pitch), Lq, B, do_pitch, kAttnBQ))) return r;
  if ((r = encode_bf16_3d(&P.tmS0, k_bf16, static_cast<uint64_t>(k_pitch), Lk, B, k_pitch, kBwdBlk))) return r;
  if ((r = encode_bf16_3d(&P.tmS1, v_bf16, static_cast<uint64_t>(v_pitch), Lk, B, v_pitch, kBwdBlk))) return r;
  P.r0_col0 = q_col0; P.r1_col0 = 0; P.s0_col0 = k_col0; P.s1_col0 = v_col0;
  P.out0 = static_cast<__nv_bfloat16*>(dq_bf16); P.out0_pitch = dq_pitch; P.out0_col0 = dq_col0;
  P.out1 = nullptr; P.out1_pitch = 0; P.out1_col0 = 0;
  {
    const long long nwork = static_cast<long long>(cdiv(Lq, kAttnItemQ)) * H * B;
    if (nwork > 0x7fffffffLL) return -3;
    seq_attn_bwd_kernel<false><<<nwork < sms ? static_cast<int>(nwork) : sms, kAttnThreads, kBwdSmemBytes, st>>>(P);
    SEQ_CUDA_OK(cudaGetLastError());
  }
  // ---- dK / dV: resident K / V tiles, streamed Q / dO blocks ----
  if ((r = encode_bf16_3d(&P.tmR0, k_bf16, static_cast<uint64_t>(k_pitch), Lk, B, k_pitch, kAttnBQ))) return r;
  if ((r = encode_bf16_3d(&P.tmR1, v_bf16, static_cast<uint64_t>(v_pitch), Lk, B, v_pitch, kAttnBQ))) return r;
  if ((r = encode_bf16_3d(&P.tmS0, q_bf16, static_cast<uint64_t>(q_pitch), Lq, B, q_pitch, kBwdBlk))) return r;
  if ((r = encode_bf16_3d(&P.tmS1, do_bf16, static_cast<uint64_t>(do_pitch), Lq, B, do_pitch, kBwdBlk))) return r;
  P.r0_col0 = k_col0; P.r1_col0 = v_col0; P.s0_col0 = q_col0; P.s1_col0 = 0;
  P.out0 = static_cast<__nv_bfloat16*>(dk_bf16); P.out0_pitch = dk_pitch; P.out0_col0 = dk_col0;
  P.out1 = static_cast<__nv_bfloat16*>(dv_bf16); P.out1_pitch = dv_pitch; P.out1_col0 = dv_col0;
  {
    const long long nwork = static_cast<long long>(cdiv(Lk, kAttnItemQ)) * H * B;
    if (nwork > 0x7fffffffLL) return -3;
    seq_attn_bwd_kernel<true><<<nwork < sms ? static_cast<int>(nwork) : sms, kAttnThreads, kBwdSmemBytes, st>>>(P);
    SEQ_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------- optimizer

static int sumsq_grid(long long n4) {
  const long long want = (n4 + 255) / 256;
  const int cap = seq_num_sms() * 4;
  return want < cap ? static_cast<int>(want > 0 ? want : 1) : cap;
}

size_t fnd_seq_grad_sumsq_workspace(long long n) {
  if (n <= 0) return 0;
  return static_cast<size_t>(sumsq_grid(n / 4)) * 4 * sizeof(float);
}

int fnd_seq_grad_sumsq(const float* g, long long n, float* out4, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g || !out4 || !workspace || n <= 0 || (n & 3)) return -1;
  if (!aligned16(g) || !aligned16(out4) || !aligned16(workspace)) return -3;
  if (workspace_bytes < fnd_seq_grad_sumsq_workspace(n)) return -4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = sumsq_grid(n / 4);
  seq_sumsq_kernel<<<grid, 256, 0, st>>>(g, n / 4, static_cast<float*>(workspace));
  SEQ_CUDA_OK(cudaGetLastError());
  seq_reduce_partials_kernel<<<1, 256, 0, st>>>(static_cast<const float*>(workspace), grid, 4, out4);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int fnd_seq_adamw_step(float* w, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, float max_norm, float grad_scale, const float* sumsq,
                       void* stream) {
  if (!w || !g || !m || !v || !sumsq || n <= 0 || (n & 3) || step < 1) return -1;
  if (!aligned16(w) || !aligned16(g) || !aligned16(m) || !aligned16(v)) return -3;
  SeqAdamParams P{w, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay,
                  1.f - powf(beta1, static_cast<float>(step)), 1.f - powf(beta2, static_cast<float>(step)), max_norm, grad_scale, sumsq};
  const long long want = (n / 4 + 255) / 256;
  const int cap = seq_num_sms() * 8;
  seq_adamw_kernel<<<want < cap ? static_cast<int>(want) : cap, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  SEQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
