/* fnd_b200.h — C ABI of libfnd_b200.so, the B200 (sm_100a) fusion hot path of Ultrafnd.
 *
 * Plain C: raw device pointers, sizes and a cudaStream_t passed as void*. No torch types.
 * Every function returns 0 on success, a negative value for a host-side error (bad argument,
 * CUDA runtime error = -1000 - cudaError_t) and a positive value for a device-side error code.
 * The caller owns every buffer; the library never allocates or frees device memory.
 *
 * All citations are relative to the reference checkout (Nuralamsiddik16/Ultrafnd_git).
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
 * Building block: C[M,N] (fp32) = A[M,K] * B[N,K]^T on the tcgen05 tensor cores.
 * Replaces ATen addmm as reached from nn.Linear (src/models/fusion/cross_modal_transformer.py:96-102,
 * 122-129; src/models/fusion/deep_truth_classifier.py:121-128). Exposed for tests and for callers that
 * want the raw GEMM; the fused entry points below are what the model uses.
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

#ifdef __cplusplus
}
#endif
#endif /* FND_B200_H_ */
