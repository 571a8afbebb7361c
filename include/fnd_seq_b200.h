/* fnd_seq_b200.h — C ABI of the sequence front-end (Tier B) inside libfnd_b200.so.
 *
 * SELF-ORACLE SCOPE: none of these operators exists in the reference (SURVEY.md §0 — no sequence attention, no
 * LayerNorm, no per-token projection in /root/reference). They implement the front-end that BASELINE.json's north_star
 * names (token / frame / audio sequences -> per-modality projection + LayerNorm -> bidirectional multi-head co-attention
 * with key-padding masks -> masked mean-pool -> the (B, D) vectors CrossModalTransformer.forward consumes,
 * src/models/fusion/cross_modal_transformer.py:141-150) and are checked against oracle/seq_oracle.py. The only
 * reference semantics restated here is the masked mean (src/core_blocks/text_blocks.py:81-86).
 *
 * Conventions as in fnd_b200.h: plain C, raw DEVICE pointers, sizes, cudaStream_t as void*; 0 = ok, negative = host-side
 * error (-1000 - cudaError_t for CUDA runtime errors). Every call is stream-ordered, allocates nothing, never
 * synchronises and can be captured into a CUDA graph (TMA descriptors travel in kernel-parameter space). bf16 matrices are
 * row-major with a row pitch in ELEMENTS (multiple of 8) and a 16-byte aligned base. `err_flag` (optional) is a device int
 * that a kernel sets to a non-zero code instead of hanging if one of its bounded waits expires.
 */
#ifndef FND_SEQ_B200_H_
#define FND_SEQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One-time, per-device kernel attribute setup (opt-in shared-memory sizes). Call once outside graph capture; every
 * other entry point also calls it lazily. */
int fnd_seq_init(void);
/* Number of CTA pairs the 256 x 256 two-SM projection kernel (csrc/fnd_seq_gemm2.cuh) keeps resident (74 on a B200);
 * 0 = that kernel is disabled (FND_SEQ_GEMM_PAIR=0) and fnd_seq_linear always takes the single-CTA kernel. */
int fnd_seq_pair_clusters(void);

/* y[i] = bf16(x[i]), n a multiple of 8 (feature tensors arrive as fp32: forensic_trainer.py:254-258 casts to fp32). */
int fnd_seq_cast_bf16(const float* x, void* y_bf16, long long n, void* stream);

/* out[M,N] = a[M,K] . w[N,K]^T (+ bias[N]) (+ resid[M,N]) (-> exact-erf GELU when act == 1), fp32 accumulation on the
 * tcgen05 tensor cores (persistent 128 x {64,128,256} tiles, two TMEM accumulators). nn.Linear over tokens: w is the
 * [out_features, in_features] weight as bf16. K and N multiples of 8. Either output may be NULL. */
int fnd_seq_linear(const void* a_bf16, int a_pitch, const void* w_bf16, int w_pitch, const float* bias,
                   const void* resid_bf16, int resid_pitch, int act, void* out_bf16, int out_pitch, float* out_f32,
                   int f32_pitch, int M, int N, int K, int* err_flag, void* stream);

/* y = LayerNorm(x + resid) over the last dimension (resid may be NULL: plain LayerNorm; biased variance, eps inside the
 * sqrt, affine gamma / beta in fp32). d a multiple of 8, d <= 2048. */
int fnd_seq_layernorm(const void* x_bf16, int x_pitch, const void* resid_bf16, int resid_pitch, const float* gamma,
                      const float* beta, float eps, void* y_bf16, int y_pitch, int M, int d, void* stream);

/* Multi-head cross-attention forward, head dimension 64:
 *   out[b,q,h,:] = softmax_k(Q[b,q,h,:] . K[b,k,h,:] * scale + key_padding_mask[b,k]) V[b,k,h,:]
 * q / k / v are [B*Lq | B*Lk, pitch] bf16 matrices; head h occupies columns col0 + 64 h .. col0 + 64 h + 63 (so fused
 * [Q|K|V] projection outputs are addressed in place). Keys with index >= kv_len[b] (NULL: Lk) or kv_mask[b,k] == 0
 * (NULL: all valid) are excluded; a query with no valid key gets zeros. lse (optional) receives the natural-log
 * logsumexp of the scaled scores as [B, H, Lq]. */
int fnd_seq_coattn_forward(const void* q_bf16, int q_pitch, int q_col0, const void* k_bf16, int k_pitch, int k_col0,
                           const void* v_bf16, int v_pitch, int v_col0, const int* kv_len, const unsigned char* kv_mask,
                           int B, int H, int Lq, int Lk, float scale, void* out_bf16, int out_pitch, float* lse,
                           int* err_flag, void* stream);

/* out[b,:] = sum_l x[b,l,:] m[b,l] / max(sum_l m[b,l], 1e-6) with m = (l < len[b]) & mask[b,l] (either may be NULL):
 * src/core_blocks/text_blocks.py:81-86. d a multiple of 8. Either output may be NULL. */
int fnd_seq_masked_mean_pool(const void* x_bf16, int x_pitch, const unsigned char* mask, const int* len, int B, int L,
                             int d, float* out_f32, int f32_pitch, void* out_bf16, int bf_pitch, void* stream);

/* Probe aid (tools/attn_stamps.py): while `stamps` is non-NULL, fnd_seq_coattn_forward records clock64() event stamps of
 * its pipeline into stamps[(cta * 32 + step) * 16 + event] for the first 32 key-block steps of every CTA: softmax
 * warpgroup t sees S (0 + 4t), holds S in registers (1 + 4t), enters the exp2 section (2 + 4t), publishes P (3 + 4t);
 * the tile's MMA warp sees P (8 + 4t), has issued P V (9 + 4t), has issued the next Q K^T (10 + 4t). `stamps` must hold
 * 16 * 32 * (number of SMs) entries. NULL (the default) disables it: the production path pays one predicate per event.
 * FND_ATTN_DBG_NOEXP=1 additionally skips the exponentials (P = 0) to expose the pure pipeline latency. Not for use
 * around graph capture. */
int fnd_seq_debug_attn_stamps(long long* stamps);
/* Probe aid (tools/gemm2_stamps.py): device buffer of 16 int64 per CTA pair that the two-SM projection kernel fills with
 * clock64 totals — MMA warp: [0] loop, [1] waiting for a drained accumulator, [2] waiting for operands, [3] issuing,
 * [5] k-blocks; producer: [4] waiting for a free stage; epilogue: [6] loop, [7] waiting for an accumulator,
 * [8..11] per-slab phases (buffer free, TMEM load, arithmetic + staging stores, fence + store issue). NULL = off. */
int fnd_seq_debug_gemm_stamps(long long* stamps);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward pass. Activation gradients travel as bf16 matrices, parameter gradients are fp32. Workspaces are caller-owned
 * device buffers (16-byte aligned) of at least the size the matching *_workspace() query returns; nothing is allocated.
 * ---------------------------------------------------------------------------------------------------------------- */

/* Fused attention backward (head dimension 64). Given the forward's inputs, its output `o`, its logsumexp `lse`
 * ([B, H, Lq], from fnd_seq_coattn_forward) and the output gradient `d_o` ([B*Lq, do_pitch], head h at columns 64 h):
 *   P = exp(S * scale - lse),  dP = dO V^T,  dS = P o (dP - rowsum(dO o O)),
 *   dq = scale * dS K,  dk = scale * dS^T Q,  dv = P^T dO        (masked keys / rows without a valid key: exact zeros)
 * dq / dk / dv are written as bf16 at columns col0 + 64 h of their matrices ([B*Lq | B*Lk, pitch]; so a fused d[Q|K|V]
 * buffer is filled in place). Three launches: a row prologue and two tcgen05 kernels (csrc/fnd_seq_attn_bwd.cuh);
 * deterministic (no atomics). */
size_t fnd_seq_coattn_backward_workspace(int B, int H, int Lq);
int fnd_seq_coattn_backward(const void* q_bf16, int q_pitch, int q_col0, const void* k_bf16, int k_pitch, int k_col0,
                            const void* v_bf16, int v_pitch, int v_col0, const void* o_bf16, int o_pitch,
                            const void* do_bf16, int do_pitch, const float* lse, const int* kv_len,
                            const unsigned char* kv_mask, int B, int H, int Lq, int Lk, float scale, void* dq_bf16,
                            int dq_pitch, int dq_col0, void* dk_bf16, int dk_pitch, int dk_col0, void* dv_bf16,
                            int dv_pitch, int dv_col0, void* workspace, size_t workspace_bytes, int* err_flag,
                            void* stream);

/* LayerNorm backward. t = the forward INPUT of the LayerNorm (x + resid already summed), dy = upstream gradient.
 * dt (bf16) = gradient w.r.t. t; dgamma / dbeta (fp32 [d] each, ADJACENT: dbeta == dgamma + d) are overwritten.
 * Partial sums are reduced in a fixed order (deterministic). */
size_t fnd_seq_layernorm_backward_workspace(int M, int d);
int fnd_seq_layernorm_backward(const void* t_bf16, int t_pitch, const void* dy_bf16, int dy_pitch, const float* gamma,
                               float eps, void* dt_bf16, int dt_pitch, float* dgamma, float* dbeta, int M, int d,
                               void* workspace, size_t workspace_bytes, void* stream);

/* out[n] = sum_m x[m, n] (fp32; bias gradients). N a multiple of 8. Fixed-order reduction. */
size_t fnd_seq_colsum_workspace(int M, int N);
int fnd_seq_colsum(const void* x_bf16, int x_pitch, int M, int N, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Weight gradient of a token-level nn.Linear: dw[n_out, k_in] (fp32) = dy[tokens, n_out]^T x[tokens, k_in]. Both operands are
 * read token-major IN PLACE (MN-major UMMA descriptors on the persistent tcgen05 GEMM: nothing is transposed in memory); the
 * token reduction is split into ranges whose partial tiles are summed in a fixed order (deterministic). n_out, k_in
 * multiples of 8; dw dense (dw_pitch == k_in) whenever the plan splits (always, for tokens >= 2048). */
size_t fnd_seq_wgrad_workspace(int tokens, int n_out, int k_in);
int fnd_seq_wgrad(const void* dy_bf16, int dy_pitch, const void* x_bf16, int x_pitch, int tokens, int n_out, int k_in,
                  float* dw, int dw_pitch, void* workspace, size_t workspace_bytes, int* err_flag, void* stream);

/* Backward of fnd_seq_masked_mean_pool: dx[b,l,:] = m[b,l] / max(sum_l m[b,l], 1e-6) * dpooled[b,:] (bf16). */
int fnd_seq_masked_mean_pool_backward(const float* dpooled, int dp_pitch, const unsigned char* mask, const int* len, int B,
                                      int L, int d, void* dx_bf16, int dx_pitch, void* stream);

/* Optimizer step of the front-end over one flat fp32 range (parameters, gradients, Adam moments at the same offsets):
 * out4[0] = sum of squares of g (fixed-order reduction); then clip_grad_norm_(max_norm) + torch.optim.AdamW semantics
 * (decoupled weight decay, bias corrections 1 - beta^step) in ONE pass: g is read as g * grad_scale (1 / world after a
 * summing all-reduce) * min(1, max_norm / (sqrt(sumsq) * grad_scale + 1e-6)). n a multiple of 4, step >= 1, max_norm <= 0
 * disables clipping. The trainer the reference uses for its own parameters: src/training/forensic_trainer.py:173-177,292-298. */
size_t fnd_seq_grad_sumsq_workspace(long long n);
int fnd_seq_grad_sumsq(const float* g, long long n, float* out4, void* workspace, size_t workspace_bytes, void* stream);
int fnd_seq_adamw_step(float* w, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, float max_norm, float grad_scale, const float* sumsq,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FND_SEQ_B200_H_ */
