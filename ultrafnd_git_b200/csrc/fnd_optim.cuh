// fnd_optim.cuh — gradient clipping + AdamW over the flat parameter arena, and bf16 shadow maintenance.
//
// The trainable parameters of fusion + classifier live in ONE contiguous fp32 arena (padding between tensors is
// zero and stays zero under AdamW), with matching flat buffers for gradients and the two Adam moments. One
// kernel therefore replaces clip_grad_norm_ + the per-tensor Python loop of torch.optim.AdamW
// (reference: src/training/forensic_trainer.py:173-177,292-298) and, in the same pass, refreshes the bf16 (hi[,lo])
// operand copies of the GEMM weights that the tensor-core kernels read — 128-bit accesses throughout.
//
// torch.optim.AdamW semantics: p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//                              p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps),  g pre-scaled by the clip coef.
#pragma once
#include "fnd_rows.cuh"

namespace fnd {

struct AdamWParams {
  float* p; const float* g; float* m; float* v;
  size_t n;                      // elements (multiple of 4)
  __nv_bfloat16* sh_hi;          // bf16 shadow of p[0 .. n_shadow)
  __nv_bfloat16* sh_lo;          // residual plane (fp32x3 mode) or null
  size_t n_shadow;
  // one tensor whose row pitch is not TMA-legal gets a re-pitched shadow (pre.0.weight: [H, H+2] -> pitch rp_pitch)
  size_t rp_begin, rp_end;       // element range inside the arena
  int rp_cols, rp_pitch;
  __nv_bfloat16* rp_hi; __nv_bfloat16* rp_lo;
  DevState* state;
  // Fused step: the gradient norm has not been reduced yet — `slots` holds the per-CTA sums of squares left by the
  // wgrad / finalize CTAs. One warp of every CTA reduces them (same fixed order everywhere => identical coefficient),
  // the CTA takes optimizer step t = state->step + 1, and the LAST CTA to finish publishes norm / coefficient / step.
  const float* slots;            // null: state->{clip_coef, bc1, bc2} were published by an earlier kernel
  int nslots;
};

__device__ __forceinline__ void shadow_store4(const AdamWParams& a, size_t i, const float4& x) {
  if (i < a.n_shadow) {
    store_bf2(a.sh_hi, a.sh_lo, i, x.x, x.y);
    store_bf2(a.sh_hi, a.sh_lo, i + 2, x.z, x.w);
  }
  if (i + 4 > a.rp_begin && i < a.rp_end) {
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t e = i + q;
      if (e >= a.rp_begin && e < a.rp_end) {
        const size_t r = (e - a.rp_begin) / a.rp_cols, c = (e - a.rp_begin) % a.rp_cols;
        __nv_bfloat16 h, l;
        split_bf16(xs[q], h, l);
        a.rp_hi[r * a.rp_pitch + c] = h;
        if (a.rp_lo) a.rp_lo[r * a.rp_pitch + c] = l;
      }
    }
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamWParams a) {
  griddep_wait();
  griddep_launch();
  DevState* S = a.state;
  const float lr = S->lr, b1 = S->beta1, b2 = S->beta2, eps = S->eps;
  float coef = S->clip_coef, bc1 = S->bc1, bc2 = S->bc2, norm = 0.f;
  int t = S->step;
  if (a.slots) {
    // all 256 threads: one or two 128-bit loads each (one L2 round trip), fixed-order combine => identical everywhere
    __shared__ double s_part[8];
    __shared__ float s_norm;
    double part = 0.0;
    const int n4 = (a.nslots + 3) >> 2;                 // the slot buffer is zero beyond nslots
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 t4 = __ldcg(reinterpret_cast<const float4*>(a.slots) + i);
      part += (static_cast<double>(t4.x) + static_cast<double>(t4.y)) + (static_cast<double>(t4.z) + static_cast<double>(t4.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += s_part[w];
      s_norm = static_cast<float>(sqrt(tot));
    }
    __syncthreads();
    norm = s_norm;
    coef = clip_coef_of(S->max_norm, norm);
    t += 1;
    bc1 = 1.0f - powf(b1, static_cast<float>(t));
    bc2 = 1.0f - powf(b2, static_cast<float>(t));
  }
  const float decay = 1.0f - lr * S->weight_decay;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const uint64_t pol = l2_policy_evict_first();
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < a.n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x * 4) {
    float4 p = ld_f4_policy(a.p + i, pol);
    const float4 g4 = ld_f4_policy(a.g + i, pol);
    float4 m = ld_f4_policy(a.m + i, pol);
    float4 v = ld_f4_policy(a.v + i, pol);
    float* pp = &p.x; float* mp = &m.x; float* vp = &v.x; const float* gp = &g4.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g = gp[q] * coef;
      pp[q] *= decay;
      mp[q] = b1 * mp[q] + (1.0f - b1) * g;
      vp[q] = b2 * vp[q] + (1.0f - b2) * g * g;
      const float denom = sqrtf(vp[q]) * inv_sqrt_bc2 + eps;
      pp[q] -= step_size * (mp[q] / denom);
    }
    st_f4_policy(a.p + i, p, pol);
    st_f4_policy(a.m + i, m, pol);
    st_f4_policy(a.v + i, v, pol);
    shadow_store4(a, i, p);
  }
  if (a.slots) {
    // Publish the step bookkeeping once every CTA has read the old state: a CTA bumps the counter after its last read
    // of *S, so whoever completes the count knows nobody still needs the old values.
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int old = atomicAdd(&S->fin_counter, 1u);
      if (old == gridDim.x - 1) {
        S->grad_norm = norm; S->clip_coef = coef; S->step = t; S->bc1 = bc1; S->bc2 = bc2;
        S->fin_counter = 0u;
      }
    }
  }
}

// Rebuild the bf16 shadows from the fp32 master (after load_state_dict / an external optimizer step).
__global__ void __launch_bounds__(256) shadow_refresh_kernel(AdamWParams a) {
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < a.n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x * 4) {
    const float4 p = *reinterpret_cast<const float4*>(a.p + i);
    shadow_store4(a, i, p);
  }
}

// Sum of squares of a flat fp32 buffer into per-CTA slots (used after a gradient all-reduce, where the norm
// must be taken over the REDUCED gradients rather than folded into the wgrad epilogues).
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ slots) {
  griddep_wait();
  griddep_launch();
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(g + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    slots[blockIdx.x] = t;
  }
}
__device__ __forceinline__ void advance_step(DevState* S) {
  S->step += 1;
  S->bc1 = 1.0f - powf(S->beta1, static_cast<float>(S->step));
  S->bc2 = 1.0f - powf(S->beta2, static_cast<float>(S->step));
}
// Optimizer-step bookkeeping alone (when the norm was already produced by the fused step's finalize kernel).
__global__ void step_kernel(DevState* S) {
  griddep_wait();
  griddep_launch();
  if (threadIdx.x == 0 && blockIdx.x == 0) advance_step(S);
}
// slots -> gradient norm + clip coefficient (+ optimizer-step bookkeeping): the consumer of the slots when no fused
// AdamW follows (fnd_train_fwd_bwd) and after sumsq_kernel (gradients reduced across ranks).
__global__ void __launch_bounds__(32) norm_finish_kernel(const float* __restrict__ slots, int nslots, DevState* S,
                                                         int update_step) {
  griddep_wait();
  griddep_launch();
  const double ss = warp_reduce_slots(slots, nslots);
  if (threadIdx.x == 0) {
    const float norm = static_cast<float>(sqrt(ss));
    S->grad_norm = norm;
    S->clip_coef = clip_coef_of(S->max_norm, norm);
    if (update_step) advance_step(S);
  }
}

}  // namespace fnd
