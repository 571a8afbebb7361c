// Microbenchmarks of the NVLink/NVSwitch access patterns the data-parallel step can use (tools/p2p_probe.py drives it).
#include <cuda_runtime.h>
#include <stdint.h>
extern "C" {
// kind 0: ld.relaxed.sys.v4 from `nsrc` sources summed; 1: weak ld.global.nc.v4; 2: multimem.ld_reduce.add.v4.f32 from mc;
// kind 3: st.global.v4 of a local buffer to `nsrc` destinations; 4: multimem.st.v4 to mc; 5: st.v2 (8 B) to nsrc dests
}
struct Ptrs { float* p[8]; };
template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(Ptrs src, int nsrc, float* mc, float* local, size_t n4, size_t off4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t e = (off4 + i) * 4;
    if (KIND == 0 || KIND == 1) {
      float4 t[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p < nsrc) {
        if (KIND == 0) asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t[p].x), "=f"(t[p].y), "=f"(t[p].z), "=f"(t[p].w) : "l"(src.p[p] + e) : "memory");
        else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t[p].x), "=f"(t[p].y), "=f"(t[p].z), "=f"(t[p].w) : "l"(src.p[p] + e));
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p < nsrc) { acc.x += t[p].x; acc.y += t[p].y; acc.z += t[p].z; acc.w += t[p].w; }
    } else if (KIND == 2) {
      float4 t;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(mc + e) : "memory");
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    } else if (KIND == 3) {
      const float4 v = *reinterpret_cast<const float4*>(local + i * 4);
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p < nsrc) *reinterpret_cast<float4*>(src.p[p] + e) = v;
    } else if (KIND == 4) {
      const float4 v = *reinterpret_cast<const float4*>(local + i * 4);
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + e), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else if (KIND == 5) {
      const float4 v = *reinterpret_cast<const float4*>(local + i * 4);
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p < nsrc) {
        *reinterpret_cast<float2*>(src.p[p] + (off4 + i) * 2) = make_float2(v.x, v.y);
      }
    }
  }
  if (KIND <= 2 && acc.x == 12345.678f) local[0] = acc.x + acc.y + acc.z + acc.w;   // keep the loads alive
  if (KIND <= 2) {
    // realistic sink: write the sum locally
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 1; i += stride) local[4] = acc.x;
  }
}
extern "C" int p2p_probe(int kind, const unsigned long long* ptrs, int nsrc, unsigned long long mc, unsigned long long local,
                         unsigned long long n4, unsigned long long off4, int grid, void* stream) {
  Ptrs s;
  for (int i = 0; i < 8; ++i) s.p[i] = i < nsrc ? reinterpret_cast<float*>(ptrs[i]) : nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* m = reinterpret_cast<float*>(mc); float* l = reinterpret_cast<float*>(local);
  switch (kind) {
    case 0: probe_kernel<0><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    case 1: probe_kernel<1><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    case 2: probe_kernel<2><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    case 3: probe_kernel<3><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    case 4: probe_kernel<4><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    case 5: probe_kernel<5><<<grid, 256, 0, st>>>(s, nsrc, m, l, n4, off4); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}
