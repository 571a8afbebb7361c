// Micro-benchmark: cycles per 128-element softmax "exp block" (fma -> exp2 -> row sum -> bf16 pack) as a function of the
// number of warps per SM sub-partition and of the share of exponentials evaluated on the FMA pipe. Guides the warp layout
// of csrc/fnd_seq_attn.cuh. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/exp_ubench tools/ubench/exp_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../ultrafnd_git_b200/csrc/fnd_seq_attn.cuh"
using namespace fnd;

template <int kPoly, int kCols, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) k(float* out, long long* cyc, int reps, float sl, float off) {
  float s[kCols];
#pragma unroll
  for (int i = 0; i < kCols; ++i) s[i] = -0.01f * (threadIdx.x + i);
  const uint64_t sl2 = pack_f32x2(sl, sl), noff = pack_f32x2(-off, -off);
  uint32_t acc = 0;
  uint64_t ps0 = 0ull, ps1 = 0ull;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < kCols; i += 4) {
      float x0, x1, x2, x3, p0, p1, p2, p3;
      const uint64_t xa = fma_f32x2(pack_f32x2(s[i], s[i + 1]), sl2, noff);
      const uint64_t xb = fma_f32x2(pack_f32x2(s[i + 2], s[i + 3]), sl2, noff);
      if ((i & 4) ? (kPoly >= 3) : (kPoly >= 1)) exp2_poly_x2(xa, p0, p1);
      else { unpack_f32x2(xa, x0, x1); p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
      if ((i & 4) ? (kPoly >= 4) : (kPoly >= 2)) exp2_poly_x2(xb, p2, p3);
      else { unpack_f32x2(xb, x2, x3); p2 = ex2_approx(x2); p3 = ex2_approx(x3); }
      ps0 = add_f32x2(ps0, pack_f32x2(p0, p1));
      ps1 = add_f32x2(ps1, pack_f32x2(p2, p3));
      acc ^= pack_bf16x2(p0, p1) + pack_bf16x2(p2, p3);
      s[i] += p0 * 1e-9f;          // keep the loop-carried inputs live without adding real work
    }
  }
  const long long t1 = clock64();
  float a, b;
  unpack_f32x2(add_f32x2(ps0, ps1), a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + __uint_as_float(acc);
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int kPoly, int kCols, int kThreads>
void run(int reps) {
  const int warps = kThreads / 32;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  k<kPoly, kCols, kThreads><<<148, kThreads>>>(out, cyc, reps, 0.18f, 0.3f);
  k<kPoly, kCols, kThreads><<<148, kThreads>>>(out, cyc, reps, 0.18f, 0.3f);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0; for (int w = 0; w < warps; ++w) m += h[w];
  m /= warps;
  printf("cols=%3d poly=%d warps/SMSP=%d: %8.1f cycles per warp-block = %5.2f cycles per SMSP per 32-lane element = %7.1f per 128x128 tile\n", kCols, kPoly,
         warps / 4, m / reps, m / reps / (kCols * (warps / 4.0)), m / reps / (kCols * (warps / 4.0)) * 128.0);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int reps = 200;
  run<0, 128, 128>(reps); run<1, 128, 128>(reps); run<2, 128, 128>(reps); run<3, 128, 128>(reps);
  run<0, 128, 256>(reps); run<1, 128, 256>(reps); run<2, 128, 256>(reps); run<3, 128, 256>(reps);
  run<0, 64, 128>(reps); run<1, 64, 128>(reps); run<2, 64, 128>(reps);
  run<0, 64, 256>(reps); run<1, 64, 256>(reps); run<2, 64, 256>(reps); run<3, 64, 256>(reps);
  run<0, 64, 512>(reps); run<1, 64, 512>(reps); run<2, 64, 512>(reps); run<3, 64, 512>(reps);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
