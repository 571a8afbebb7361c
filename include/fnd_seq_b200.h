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

/* Probe aid (tools/seq_probe.py): while `stamps` is non-NULL, fnd_seq_coattn_forward launches an instrumented build in
 * which one softmax thread per CTA accumulates clock64() cycles per phase of its key-block loop into
 * stamps[cta * 8 + phase] (0 wait for S, 1 TMEM load, 2 mask + row max, 3 exp2 + pack, 4 P store + fence, 5 previous
 * P V product, 6 rescale, 7 loop / item epilogue). `stamps` must hold 8 * 2 * (number of SMs) entries. NULL restores the
 * production kernel. Not for use around graph capture. */
int fnd_seq_debug_attn_stamps(long long* stamps);

#ifdef __cplusplus
}
#endif
#endif /* FND_SEQ_B200_H_ */
