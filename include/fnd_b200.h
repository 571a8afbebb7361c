/* fnd_b200.h — C ABI of libfnd_b200.so, the B200 (sm_100a) fusion hot path of Ultrafnd.
 *
 * Plain C: raw device pointers, sizes and a cudaStream_t passed as void*. No torch types.
 * Every int-returning function returns 0 on success, a negative value for a host-side error (bad argument;
 * CUDA runtime error = -1000 - cudaError_t) and a positive value for a device-side error code.
 * The caller owns every DEVICE buffer (parameters, gradients, optimizer state, bf16 shadows, workspace);
 * the library never allocates or frees device memory. A plan handle is a small host object.
 * All hot-path entry points are stream-ordered, never synchronise and are CUDA-graph capturable.
 *
 * Citations are relative to the reference checkout (Nuralamsiddik16/Ultrafnd_git); each entry point names the
 * reference code it replaces. The reference has no native/FFI layer (it is pure Python on ATen), so these are
 * the entry points a binding at the nn.Module.forward / trainer-step boundary needs.
 */
#ifndef FND_B200_H_
#define FND_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library version (major*100 + minor) and the single architecture it is built for. */
int fnd_version(void);
const char* fnd_build_arch(void);

/* ---------------------------------------------------------------------------------------------
 * Model dimensions. Mirrors configs/model_configs/fusion.yaml + classifier.yaml and the input widths
 * hard-coded at cross_modal_transformer.py:96-99.
 * ------------------------------------------------------------------------------------------- */
typedef struct fnd_dims {
  int hidden;            /* H: 512 or 1024 (fusion hidden == classifier hidden == classifier input_dim) */
  int d_text, d_audio, d_visual, d_temporal, d_gnn;   /* 768, 128, 512, 256, 128 (multiples of 64) */
  int use_gnn;           /* 1: 16 slots in fused_cat, 0: 15 */
  int aux_dim;           /* 2 or 0 */
  int trees, depth;      /* NODE ensemble: trees*depth <= 32, depth <= 4 */
  float fusion_dropout;  /* fusion.yaml:3 */
  float clf_dropout;     /* classifier.yaml:4 */
  float tree_dropout;    /* 0.3, deep_truth_classifier.py:134 */
  float node_tau;        /* 10.0 */
} fnd_dims;

/* ---------------------------------------------------------------------------------------------
 * Parameter arena. All parameters of CrossModalTransformer + DeepTruthClassifier live in one flat fp32
 * buffer; entry i describes one state_dict tensor: its name prefixed "fusion." or "clf.", its element
 * offset, its shape (ndim 0..2) and whether it is "hot" (receives a gradient in the reference training step,
 * forensic_trainer.py:286-298). Hot tensors occupy [0, fnd_arena_hot_elems); GEMM weights come first and have
 * bf16 shadows at the same element offsets in [0, fnd_arena_shadow_elems) (the shadow buffers must hold
 * fnd_arena_shadow_buffer_elems elements: the tail is a re-pitched copy of pre.0.weight).
 * ------------------------------------------------------------------------------------------- */
int fnd_param_count(const fnd_dims* dims);
int fnd_param_info(const fnd_dims* dims, int index, char* name, int name_cap, long long* offset, int* ndim,
                   int* rows, int* cols, int* hot);
long long fnd_arena_total_elems(const fnd_dims* dims);
long long fnd_arena_hot_elems(const fnd_dims* dims);
long long fnd_arena_shadow_elems(const fnd_dims* dims);
long long fnd_arena_shadow_buffer_elems(const fnd_dims* dims);

/* ---------------------------------------------------------------------------------------------
 * Plans. A plan fixes (dims, batch, precision mode) and owns the TMA descriptors and kernel tables.
 *   mode 0: bf16 operands, fp32 accumulate.   mode 1: "fp32x3" — bf16 hi/lo operand pairs, three tensor-core
 *   products per GEMM, fp32-equivalent results (rel. err ~1e-5).
 * fnd_plan_bind uploads the tables; it is the only plan call that issues copies and must not be captured.
 * shadow_lo may be NULL in mode 0. m / v may be NULL for inference-only plans.
 * ------------------------------------------------------------------------------------------- */
int fnd_plan_create(const fnd_dims* dims, int batch, int mode, void** plan_out);
void fnd_plan_destroy(void* plan);
size_t fnd_plan_workspace_bytes(const void* plan);
int fnd_plan_bind(void* plan, void* workspace, float* params, float* grads, float* adam_m, float* adam_v,
                  void* shadow_hi, void* shadow_lo, void* stream);
/* Optional, BEFORE fnd_plan_bind: a bf16 buffer of fnd_arena_hot_elems elements that the weight-gradient kernels fill
 * with bf16 copies of the GEMM-weight gradients (same element offsets as the gradient arena; pre.0.weight excluded).
 * The data-parallel step reduces this mirror through the NVSwitch (fnd_dp_bind). NULL = no mirror. 256-byte aligned. */
int fnd_plan_set_grad_mirror(void* plan, void* grads_bf16);
/* Byte offset / element count of a named workspace buffer ("fused", "fusion_logits", "rowstat", "logits",
 * "probs", "loss_row", "dlogits", "dfused", "state", "fused_cat_hi", ...); returns -1 if unknown. */
long long fnd_plan_buffer_offset(const void* plan, const char* name);
long long fnd_plan_buffer_bytes(const void* plan, const char* name);

/* Inputs of one batch. x[i]: fp32 [*, d_i] with row pitch pitch[i] (elements, multiple of 4), i = text, audio,
 * visual, temporal, gnn. gather (optional): int64 row indices applied to x[*], aux and labels, so a whole
 * feature cache can stay resident on the device (replaces CachedTensorDataset + default collate +
 * gnn_Z[global_idx], forensic_trainer.py:60-83,238-263). aux: fp32 [*,2]; labels: int64. */
typedef struct fnd_inputs {
  const float* x[5];
  int pitch[5];
  const float* aux;
  int aux_pitch;
  const long long* labels;
  const long long* gather;
} fnd_inputs;

/* ---- optimizer / RNG state (small stream-ordered writes; call outside graph capture) ---- */
int fnd_set_hyper(void* plan, float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                  void* stream);
int fnd_set_lr(void* plan, float lr, void* stream);
int fnd_set_seed(void* plan, unsigned long long seed, void* stream);
int fnd_set_loss_scale(void* plan, float scale, void* stream);   /* d(mean loss)/d(row loss); default 1/batch */
/* Per-step loss without a copy: `host_mapped` = pinned, device-accessible HOST memory of `ring` floats (cudaHostAlloc /
 * torch pin_memory under unified addressing). Every fnd_train_step / fnd_train_fwd_bwd then ALSO stores its mean loss
 * (forensic_trainer.py:287, the value loss.item() would return at :301) into host_mapped[steps_taken % ring] from the CTA
 * that computes it — a 4-byte device-to-host store instead of a stream-ordered D2H memcpy between two steps. Readable on
 * the host once the step has completed (event / stream synchronisation). NULL switches it off. Bound plans only. */
int fnd_set_loss_mirror(void* plan, float* host_mapped, int ring, void* stream);
/* Epoch bookkeeping (ForensicTrainer._epoch_loop, forensic_trainer.py:301-313: y, p1, row losses and the three forensic
 * scalars of every batch): appends the first k rows of the LAST step's results ("loss_row", "probs"[:,1], the gathered labels,
 * "rowstat"[:, :3]) at row `offset` of caller-owned device buffers, and `batch_no` into bid — one stream-ordered launch
 * instead of five indexing ops per step; any destination may be NULL. forensic3 is [*, 3] row-major. */
int fnd_collect_rows(void* plan, int k, long long offset, long long batch_no, float* loss_rows, float* p1, long long* ys,
                     long long* bid, float* forensic3, void* stream);
/* Rebuild the bf16 operand copies from the fp32 master parameters (after load_state_dict or an external
 * optimizer step). fnd_clip_adamw_step keeps them current by itself. */
int fnd_refresh_shadows(void* plan, void* stream);

/* ---- module-level entry points (what a binding of the two nn.Modules calls) ----
 * fnd_fusion_forward  : CrossModalTransformer.forward, cross_modal_transformer.py:134-210.
 *                       Results in workspace buffers "fused" [B,H], "fusion_logits" [B,2], "rowstat" [B,16]
 *                       (cols 0..2 = semantic_conflict, emotion_intensity, temporal_delay).
 * fnd_classifier_forward : DeepTruthClassifier.forward, deep_truth_classifier.py:148-171. `fused` NULL = use
 *                       the plan's own "fused"; aux NULL = aux already staged by fnd_fusion_forward's inputs.
 *                       Results in "logits", "probs" [B,2].
 * fnd_ce_loss_fwd_bwd : F.cross_entropy(logits, y) mean + its gradient, forensic_trainer.py:287.
 *                       Results in "loss_row" [B], "dlogits" [B,2].
 * fnd_classifier_backward / fnd_fusion_backward : autograd of the two forwards; parameter gradients land in
 *                       the bound gradient arena, "dfused" [B,H] carries the gradient across the boundary.
 *                       dlogits / dfused NULL = use the workspace buffers; dfusion_logits may be NULL.
 * fnd_clip_adamw_step : clip_grad_norm_(max_norm) + AdamW.step, forensic_trainer.py:292-298. norm_from_slots
 *                       = 1 reuses the norm folded into the fused step's wgrad epilogues; 0 recomputes it over
 *                       the gradient arena (use after a gradient all-reduce or the module-level backward). */
int fnd_fusion_forward(void* plan, const fnd_inputs* in, int training, void* stream);
int fnd_classifier_forward(void* plan, const float* fused, const float* aux, int aux_pitch, int training,
                           void* stream);
int fnd_ce_loss_fwd_bwd(void* plan, const long long* labels, void* stream);
int fnd_classifier_backward(void* plan, const float* dlogits, void* stream);
int fnd_fusion_backward(void* plan, const float* dfused, const float* dfusion_logits, void* stream);
int fnd_clip_adamw_step(void* plan, int norm_from_slots, void* stream);

/* ---- fused trainer-step entry points (ForensicTrainer._epoch_loop body, forensic_trainer.py:285-298) ----
 * fnd_train_fwd_bwd : forward + cross-entropy + full backward in one stream-ordered sequence; leaves the mean
 *                     loss / gradient norm in "state" and all gradients in the arena.
 * fnd_train_step    : fnd_train_fwd_bwd followed by fnd_clip_adamw_step(norm_from_slots = 1).
 * fnd_train_step_overlap : fnd_train_step with the fuse_mlp.0 / fuse_mlp.3 weight gradients (70 % of the gradient
 *                     bytes) launched on `side_stream` as soon as dgrad_fuse1 has produced their operands, so that
 *                     their HBM write-back runs UNDER the latency-bound rest of the backward chain; the streams fork
 *                     and join through events (capturable into one CUDA graph). Results are bit-identical to
 *                     fnd_train_step (same tiles, same norm slots). side_stream NULL = fnd_train_step.
 * fnd_eval_step     : forward of both modules (+ row losses when labels are given), no dropout, nothing saved. */
int fnd_train_fwd_bwd(void* plan, const fnd_inputs* in, void* stream);
int fnd_train_step(void* plan, const fnd_inputs* in, void* stream);
int fnd_train_step_overlap(void* plan, const fnd_inputs* in, void* stream, void* side_stream);
int fnd_eval_step(void* plan, const fnd_inputs* in, void* stream);

/* ---- data-parallel optimizer step over NVLink peer memory (batch-sharded replicas, one process per GPU) ----
 * Replaces, for N ranks, "all-reduce the gradients, then clip_grad_norm_ + AdamW on every rank"
 * (forensic_trainer.py:292-298 under a data-parallel wrapper). Every rank owns 1/N of the arena (its slice, three
 * element ranges: fnd_dp_shard_ranges). Each rank stores the parts of its gradient that its peers own into their
 * staging buffers (P2P stores over NVLink; bf16 on the wire when stage_bf16 = 1, summed in fp32), the owner adds the
 * N pieces in rank order, the clip coefficient is formed from the N slice norms in rank order (bit-identical on all
 * ranks), AdamW runs on the slice only (fp32 master / m / v stay sharded) and the refreshed bf16 operand shadows + the
 * small fp32 parameters are stored into every rank's buffers.
 * Every rank places params | grads | shadow_hi | shadow_lo | a 256-byte zeroed comm pad | a staging region
 * (fnd_dp_stage_bytes) at the SAME byte offsets of one peer-mapped (symmetric) allocation; peer_bases[p] is rank p's
 * base address as mapped into THIS process; multicast_base (0 = none) is the NVSwitch multicast mapping of the same
 * allocation — when given, the all-gather uses one multimem.st per 16 bytes instead of one store per peer, and with
 * off_grads_bf16 >= -1 the reduce-scatter becomes ONE multimem.ld_reduce kernel (in-switch sum over all gradient arenas;
 * off_grads_bf16 >= 0 is the offset of the bf16 gradient mirror given to fnd_plan_set_grad_mirror, read instead of the
 * fp32 arena for the GEMM weights; -2 keeps the store-based exchange through the staging region). The plan must already be bound (fnd_plan_bind) to this rank's buffers.
 * gred: local scratch of fnd_dp_stage_bytes / (world * elem size) floats; slots: 1024 zeroed floats.
 * fnd_dp_optimizer_step is stream-ordered and graph-capturable; every rank must call it once per fnd_train_fwd_bwd.
 * After it, only the OWNER of a slice holds current fp32 master weights for it; gather the slices (e.g. one broadcast
 * per range) before reading a state_dict. */
int fnd_dp_bind(void* plan, int rank, int world, const unsigned long long* peer_bases, long long off_params,
                long long off_grads, long long off_shadow_hi, long long off_shadow_lo, long long off_pad,
                long long off_stage, int stage_bf16, unsigned long long multicast_base, long long off_grads_bf16,
                float* gred, long long gred_elems, float* slots, long long slots_elems);
long long fnd_dp_stage_bytes(const void* plan, int world, int stage_bf16);
/* The slice of [0, hot) rank `rank` owns: three element ranges (its shares of fuse_mlp.0/.3 weights and of the arena
 * before / after them); returns the number of ranges written to lo3 / hi3. */
int fnd_dp_shard_ranges(const void* plan, int rank, int world, long long* lo3, long long* hi3);
int fnd_dp_optimizer_step(void* plan, void* stream);
/* fnd_train_fwd_bwd + fnd_dp_optimizer_step as one call. With a non-NULL side_stream the fuse_mlp.0 / fuse_mlp.3 weight
 * gradients (70 % of the gradient bytes) are produced early and pushed to their owners from side_stream UNDER the rest
 * of the backward pass (event fork / join; capturable as one graph). */
int fnd_train_step_dp(void* plan, const fnd_inputs* in, void* stream, void* side_stream, int flags);
/* flags (need side_stream): 1 = early push as described above; 2 = DEFER the optimizer update + all-gather of the
 * fuse_mlp.0/.3 slice (70 % of the all-gather bytes): the next fnd_train_step_dp applies it from side_stream under its
 * first four kernels and joins before gemm_fuse0. While an update is pending the fuse_mlp shadows are one step old:
 * call fnd_dp_flush (every rank) before anything else reads the model — an evaluation pass, a learning-rate or
 * hyper-parameter change, gathering the master weights. fnd_dp_optimizer_step flushes by itself. */
int fnd_dp_flush(void* plan, void* stream);

/* Per-kernel timing for benchmarks: between begin and end every kernel launch of this plan is followed by a
 * cudaEvent on `stream`; end synchronises and returns, per kernel name (64-byte slots in `names`), the summed
 * milliseconds between consecutive events. Not capturable; do not use around graph replays. */
int fnd_profile_begin(void* plan, void* stream);
int fnd_profile_end(void* plan, void* stream, char* names, float* ms, int cap, int* count);
/* Debug aid: limit >= 0 makes every entry point issue only its first `limit` kernel launches (prefix timing of the
 * step, tools/prefix_probe.py); -1 restores normal operation. Results are incomplete while a limit is set. */
int fnd_debug_set_launch_limit(void* plan, int limit);
/* Test aid: 1 (default) = split-K launches whose problems share 2 / 4 / 8 splits run as thread-block clusters and exchange
 * their partial tiles through distributed shared memory; 0 = through the L2 workspace (round-1 path). Affects launches
 * issued AFTER the call (plans keep their tile / split choices); both exchanges sum in the same split order, so results are
 * bit-identical. Returns the previous value. Call outside graph capture; captured graphs keep what they captured. */
int fnd_debug_set_cluster_splitk(int on);
/* Number of kernel launches one call of the named entry point issues ("train_step", "eval_step", ...). */
int fnd_launch_count(const void* plan, const char* entry);
/* Dropout keep-multipliers (0 or 1/(1-p)) that the NEXT training forward will use for a layer
 * (1 fuse0 [B,2H], 2 fuse1 [B,H], 3 pre0 [B,H], 4 pre1 [B,H], 5 tree [B,T,2]); test support. */
int fnd_export_dropout_mask(void* plan, int layer, float* out, long long n, void* stream);
/* Reads (and clears) the device error flag; synchronises the stream. */
int fnd_check_error(void* plan, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Building block: C[M,N] (fp32) = A[M,K] * B[N,K]^T on the tcgen05 tensor cores.
 * Replaces ATen addmm as reached from nn.Linear (cross_modal_transformer.py:96-102,122-129;
 * deep_truth_classifier.py:121-128). Exposed for tests and for callers that want the raw GEMM.
 *   a_mn / b_mn : 0 = operand stored [rows][K] (K contiguous), 1 = stored [K][rows] (rows contiguous)
 *   *_lo        : residual bf16 planes, required when ncombo == 3 (fp32x3 mode), else may be NULL
 *   bn          : tile width (32, 64 or 128; MN-major B needs >= 64);  splits : split-K factor
 *   scratch     : device buffer of fnd_gemm_scratch_bytes() bytes, 256-byte aligned
 * Synchronises the stream before returning (utility, not a hot-path call).
 * ------------------------------------------------------------------------------------------- */
size_t fnd_gemm_scratch_bytes(int M, int N, int bn, int splits);
int fnd_gemm_bf16(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                  int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                  void* scratch, size_t scratch_bytes, void* stream);

/* Stream-ordered variant (plain bf16 operands, ncombo = 1): no synchronisation, no host read-back; a device-side
 * timeout is reported through `err_flag` (device int, optional — else the flag inside `scratch`). `scratch` must stay
 * untouched until the launch has completed. Used for the weight-gradient GEMMs of the sequence front-end
 * (dW[N,K] = dY[M,N]^T X[M,K]: a_mn = b_mn = 1, nothing is transposed in memory). */
int fnd_gemm_bf16_async(const void* a_hi, int a_pitch, int a_mn, const void* b_hi, int b_pitch, int b_mn, float* c,
                        int c_pitch, int M, int N, int K, int bn, int splits, void* scratch, size_t scratch_bytes,
                        int* err_flag, void* stream);

/* Probe variant: launches the kernel `reps` times back to back and, when `stamps` is non-NULL, has every CTA record
 * eight clock64() stamps into stamps[cta*8 + i] (0 start, 1 setup done, 2 first operands landed, 3 last MMA issued,
 * 4 accumulator ready, 5 split-K exchange done, 6 epilogue done, 7 all warps done). Used by tools/gemm_probe.py. */
int fnd_gemm_bf16_probe(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                        int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits,
                        int ncombo, void* scratch, size_t scratch_bytes, void* stream, long long* stamps, int reps);

#ifdef __cplusplus
}
#endif
#endif /* FND_B200_H_ */
