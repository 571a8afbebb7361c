// fnd_api.cu — C-ABI of libfnd_b200.so (see include/fnd_b200.h for the contract of every entry point).
#include "../../include/fnd_b200.h"
#include "fnd_gemm_host.h"

using namespace fnd;

#define FND_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return -1000 - static_cast<int>(_e); \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" {

int fnd_version(void) { return 100; }

const char* fnd_build_arch(void) { return "sm_100a"; }

size_t fnd_gemm_scratch_bytes(int M, int N, int bn, int splits) {
  size_t b = align_up(sizeof(GemmProblem), 256);
  b += 256;                                                              // error flag
  b += align_up(sizeof(int) * ceil_div(M, kGemmBM) * ceil_div(N, bn), 256);   // split-K counters
  if (splits > 1) b += splitk_ws_floats(M, N, bn, splits) * sizeof(float);
  return b + 256;
}

int fnd_gemm_bf16(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                  int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                  void* scratch, size_t scratch_bytes, void* stream) {
  if (!a_hi || !b_hi || !c || !scratch) return -1;
  if (scratch_bytes < fnd_gemm_scratch_bytes(M, N, bn, splits)) return -2;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255) != 0) return -3;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  GemmProblem* dtab = reinterpret_cast<GemmProblem*>(base);
  size_t off = align_up(sizeof(GemmProblem), 256);
  int* derr = reinterpret_cast<int*>(base + off);
  off += 256;
  int* dctr = reinterpret_cast<int*>(base + off);
  const size_t ctr_bytes = align_up(sizeof(int) * ceil_div(M, kGemmBM) * ceil_div(N, bn), 256);
  off += ctr_bytes;
  float* dws = reinterpret_cast<float*>(base + off);

  EpiParams epi;
  memset(&epi, 0, sizeof(epi));
  epi.out_f32 = c;
  epi.f32_pitch = c_pitch;
  Operand A{static_cast<const __nv_bfloat16*>(a_hi), static_cast<const __nv_bfloat16*>(a_lo), a_pitch, a_mn != 0};
  Operand B{static_cast<const __nv_bfloat16*>(b_hi), static_cast<const __nv_bfloat16*>(b_lo), b_pitch, b_mn != 0};
  GemmProblem hp;
  int r = fill_problem(hp, A, B, M, N, K, bn, splits, ncombo, kEvictNormal, kEvictNormal, dws, dctr, epi);
  if (r) return r;
  const int grid = finish_table(&hp, 1);
  FND_CUDA_OK(cudaMemsetAsync(derr, 0, 256 + ctr_bytes, st));
  FND_CUDA_OK(cudaMemcpyAsync(dtab, &hp, sizeof(hp), cudaMemcpyHostToDevice, st));
  RunCtx ctx{derr, nullptr};
  const int kind = (a_mn ? (b_mn ? 2 : 3) : (b_mn ? 1 : 0));
  FND_CUDA_OK(launch_gemm(kind, dtab, 1, grid, ctx, st));
  int herr = 0;
  FND_CUDA_OK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));
  return herr;
}

}  // extern "C"
