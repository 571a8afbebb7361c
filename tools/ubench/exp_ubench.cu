// Micro-benchmark: sustained MUFU.EX2 and FMA-pipe exp2 throughput per SM sub-partition as a function of resident warps.
// Every exponential feeds the next iteration (8 independent chains per thread), so nothing can be hoisted or folded.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/exp_ubench tools/ubench/exp_ubench.cu
// Result on a B200 (profiles/r02_exp_ubench.txt): MUFU.EX2 retires one warp instruction per 8.0 cycles per sub-partition
// (16 exp / clk / SM) once two warps feed it, a lone warp gets one per 10.8; in the softmax mix (fma + exp2 + add + bf16
// pack) 8.8 - 9.5; the cubic polynomial on the FMA pipe costs 9.0 - 9.5 — the same as the MUFU it would relieve.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../ultrafnd_git_b200/csrc/fnd_seq_attn.cuh"
using namespace fnd;

template <int kMode>   // 0: MUFU only, 1: MUFU + fma + add + bf16 pack (the softmax mix), 2: polynomial only
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int reps) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
  uint32_t acc = 0;
  float sum = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float p0, p1;
      if (kMode == 2) {
        exp2_poly_x2(pack_f32x2(x[i], x[i + 1]), p0, p1);
      } else if (kMode == 1) {
        float a, b;
        unpack_f32x2(fma_f32x2(pack_f32x2(x[i], x[i + 1]), pack_f32x2(0.5f, 0.5f), pack_f32x2(-0.25f, -0.25f)), a, b);
        p0 = ex2_approx(a); p1 = ex2_approx(b);
        sum += p0 + p1;
        acc ^= pack_bf16x2(p0, p1);
      } else {
        p0 = ex2_approx(x[i]); p1 = ex2_approx(x[i + 1]);
      }
      x[i] = p0 - 1.0f; x[i + 1] = p1 - 1.0f;      // stays in [-1, 0]: the chain never saturates
    }
  }
  const long long t1 = clock64();
  float s = sum + __uint_as_float(acc & 0xffu);
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int kMode>
void run(int warps, int reps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  k<kMode><<<148, warps * 32>>>(out, cyc, reps);
  k<kMode><<<148, warps * 32>>>(out, cyc, reps);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0; for (int w = 0; w < warps; ++w) m += h[w];
  m /= warps;
  const double per_smsp = m / (8.0 * reps) / (warps / 4.0);      // cycles per warp-wide exponential per sub-partition
  printf("mode %d warps/SMSP=%d: %6.2f cycles per warp-exponential per SMSP  (%5.1f exp/clk/SM)\n", kMode, warps / 4, per_smsp, 4 * 32 / per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int reps = 2000;
  for (int w : {4, 8, 16}) run<0>(w, reps);
  for (int w : {4, 8, 16}) run<1>(w, reps);
  for (int w : {4, 8, 16}) run<2>(w, reps);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
