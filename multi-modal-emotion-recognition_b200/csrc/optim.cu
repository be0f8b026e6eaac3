// Flat fused Adam (coupled L2 weight decay, bias correction) with optional global-norm
// gradient clipping and a bf16 shadow copy of the updated weights for the tensor-core
// GEMMs.  28 B/param of fp32 traffic (+2 B for the shadow): a pure HBM-bandwidth kernel.
#include <stdlib.h>

#include "common.cuh"

namespace mmer {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ shadow, long long n, float lr_over_bc1, float beta1, float beta2, float omb1, float omb2, float eps,
            float wd, float inv_sqrt_bc2, float grad_scale, const float* __restrict__ sumsq, float max_norm) {
  float gs = grad_scale;
  if (sumsq != nullptr) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(*sumsq) * fabsf(grad_scale);
    gs *= fminf(max_norm / (total + 1e-6f), 1.f);
  }
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
    float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = fmaf(wd, pp[k], gg[k] * gs);
      mm[k] = fmaf(beta1, mm[k], omb1 * gr);
      vq[k] = fmaf(beta2, vq[k], omb2 * gr * gr);
      pp[k] -= lr_over_bc1 * mm[k] / (sqrtf(vq[k]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (shadow != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + i) = pk;
    }
  } else {
    for (long long k = i; k < n; ++k) {
      const float gr = fmaf(wd, p[k], g[k] * gs);
      const float mk = fmaf(beta1, m[k], omb1 * gr);
      const float vk = fmaf(beta2, v[k], omb2 * gr * gr);
      m[k] = mk; v[k] = vk;
      p[k] -= lr_over_bc1 * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
      if (shadow != nullptr) shadow[k] = __float2bfloat16_rn(p[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Data-parallel step: gradient reduce-scatter + Adam on this rank's shard + parameter all-gather in ONE kernel over
// NVSwitch multicast (NVLS) addresses.  g_mc / p_mc / shadow_mc are multicast pointers of symmetric buffers:
//   multimem.ld_reduce  -- the switch returns the SUM over all ranks of the gradient words (no NCCL all-reduce);
//   multimem.st         -- the updated fp32 weights and their bf16 shadow are written to every rank at once.
// Each element is reduced and updated by exactly one rank, so all ranks hold bit-identical parameters by
// construction; Adam's moments exist only for the rank's own shard.  The caller brackets the launch with a
// cross-rank barrier on both sides (gradients complete before, parameters landed after).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st_b32x2(void* mc, uint32_t a, uint32_t b) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(a)),
               "f"(__uint_as_float(b))
               : "memory");
}

// clip_grad_norm_ in the multicast step: sum of squares of the REDUCED gradient over this rank's shard ...
__global__ void __launch_bounds__(256)
sumsq_shard_mc_kernel(const float* __restrict__ g_mc, long long lo, long long hi, float* __restrict__ out) {
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += stride) {
    const float4 v = mc_ld_reduce_add(g_mc + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ float sw[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sw[k];
    atomicAdd(out, t);
  }
}
// ... published into slot `rank` of a symmetric array on every rank; after a barrier each rank adds the slots in the
// same order, so all ranks clip by the identical factor
__global__ void publish_slot_mc_kernel(const float* __restrict__ local, float* __restrict__ slots_mc, int rank) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(slots_mc + rank), "f"(*local) : "memory");
}

// Four independent 16-byte chunks per thread: the switch-reduced gradient load (multimem.ld_reduce) is a ~2-4 us round
// trip through NVSwitch, and with one chunk per thread the kernel was bound by that latency (0.12 ms for a 1/8 shard of
// 7.77 M parameters, profiles/r02_bench_n8.json dp_wait) -- all four reductions and the twelve local loads are issued
// before the first result is used.
constexpr int ADAM_MC_UNROLL = 4;
__global__ void __launch_bounds__(256)
adam_shard_mc_kernel(const float* __restrict__ p, float* __restrict__ p_mc, const float* __restrict__ g_mc,
                     float* __restrict__ m, float* __restrict__ v, bf16* __restrict__ shadow_mc, long long lo, long long hi,
                     float lr_over_bc1, float beta1, float beta2, float omb1, float omb2, float eps, float wd,
                     float inv_sqrt_bc2, float grad_scale, const float* __restrict__ sumsq_slots, int n_slots, float max_norm) {
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long i0 = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;   // lo, hi, buffer length: multiples of 4
  if (i0 >= hi) return;
  if (sumsq_slots != nullptr) {
    // torch.nn.utils.clip_grad_norm_ on the averaged gradient: coef = max_norm / (total_norm + 1e-6), clamped to 1
    float ss = 0.f;
    for (int k = 0; k < n_slots; ++k) ss += sumsq_slots[k];
    const float total = sqrtf(ss) * fabsf(grad_scale);
    grad_scale *= fminf(max_norm / (total + 1e-6f), 1.f);
  }
  float4 gv[ADAM_MC_UNROLL], pv[ADAM_MC_UNROLL], mv[ADAM_MC_UNROLL], vv[ADAM_MC_UNROLL];
#pragma unroll
  for (int u = 0; u < ADAM_MC_UNROLL; ++u) {
    const long long i = i0 + (long long)u * nthreads * 4;
    if (i < hi) gv[u] = mc_ld_reduce_add(g_mc + i);
  }
#pragma unroll
  for (int u = 0; u < ADAM_MC_UNROLL; ++u) {
    const long long i = i0 + (long long)u * nthreads * 4;
    if (i < hi) {
      pv[u] = *reinterpret_cast<const float4*>(p + i);
      mv[u] = *reinterpret_cast<float4*>(m + (i - lo));   // the moments exist for this rank's shard only (ZeRO-1)
      vv[u] = *reinterpret_cast<float4*>(v + (i - lo));
    }
  }
#pragma unroll
  for (int u = 0; u < ADAM_MC_UNROLL; ++u) {
    const long long i = i0 + (long long)u * nthreads * 4;
    if (i >= hi) continue;
    float* pp = &pv[u].x; const float* gg = &gv[u].x; float* mm = &mv[u].x; float* vq = &vv[u].x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // same arithmetic, in the same order, as adam_kernel
      const float gr = fmaf(wd, pp[k], gg[k] * grad_scale);
      mm[k] = fmaf(beta1, mm[k], omb1 * gr);
      vq[k] = fmaf(beta2, vq[k], omb2 * gr * gr);
      pp[k] -= lr_over_bc1 * mm[k] / (sqrtf(vq[k]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(m + (i - lo)) = mv[u];
    *reinterpret_cast<float4*>(v + (i - lo)) = vv[u];
    mc_st_f32x4(p_mc + i, pv[u]);
    if (shadow_mc != nullptr) {
      __nv_bfloat162 l2 = __floats2bfloat162_rn(pv[u].x, pv[u].y), h2 = __floats2bfloat162_rn(pv[u].z, pv[u].w);
      mc_st_b32x2(shadow_mc + i, *reinterpret_cast<uint32_t*>(&l2), *reinterpret_cast<uint32_t*>(&h2));
    }
  }
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long k = i; k < n; ++k) s += g[k] * g[k];
    }
  }
  __shared__ float sw[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sw[k];
    atomicAdd(out, t);
  }
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int64_t step, float grad_scale, const float* sumsq,
                   float max_norm, void* stream) {
  MMER_CHECK_ARG(p && g && m && v, "adam: null pointer");
  MMER_CHECK_ARG(step >= 1, "adam: step counts from 1");
  MMER_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                 "adam: buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  // The reference's optimizer holds the betas as Python doubles (0.9, 0.999) and derives 1-beta and the bias
  // corrections in double; recover the decimal the caller meant from the float that crossed the C ABI.
  char buf[32];
  snprintf(buf, sizeof(buf), "%.7g", (double)beta1);
  const double b1 = strtod(buf, nullptr);
  snprintf(buf, sizeof(buf), "%.7g", (double)beta2);
  const double b2 = strtod(buf, nullptr);
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  const long long nt = (n + 3) / 4;
  adam_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, (bf16*)shadow_bf16, n, (float)(lr / bc1), beta1, beta2, (float)(1.0 - b1),
      (float)(1.0 - b2), eps, weight_decay,
      (float)(1.0 / sqrt(bc2)), grad_scale, sumsq, max_norm);
  MMER_LAUNCH_CHECK("adam_kernel");
  return 0;
}

int mmer_grad_sumsq_multicast(const float* g_mc, int64_t lo, int64_t hi, float* local_acc, float* slots_mc, int rank,
                              void* stream) {
  MMER_CHECK_ARG(g_mc && local_acc && slots_mc && rank >= 0, "grad_sumsq_multicast: bad argument");
  MMER_CHECK_ARG(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "grad_sumsq_multicast: shard bounds must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(local_acc, 0, sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "memset(sumsq shard)");
  if (hi > lo) {
    long long want = ((hi - lo) / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    sumsq_shard_mc_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(g_mc, lo, hi, local_acc);
    MMER_LAUNCH_CHECK("sumsq_shard_mc_kernel");
  }
  publish_slot_mc_kernel<<<1, 1, 0, st>>>(local_acc, slots_mc, rank);
  MMER_LAUNCH_CHECK("publish_slot_mc_kernel");
  return 0;
}

int mmer_adam_step_multicast(const float* p_local, float* p_mc, const float* g_mc, float* m, float* v, void* shadow_mc,
                             int64_t lo, int64_t hi, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int64_t step, float grad_scale, const float* sumsq_slots, int n_slots, float max_norm,
                             void* stream) {
  MMER_CHECK_ARG(sumsq_slots == nullptr || (n_slots >= 1 && n_slots <= 64 && max_norm > 0.f),
                 "adam_multicast: clipping needs 1..64 slots and a positive max_norm");
  MMER_CHECK_ARG(p_local && p_mc && g_mc && m && v, "adam_multicast: null pointer");
  MMER_CHECK_ARG(step >= 1, "adam_multicast: step counts from 1");
  MMER_CHECK_ARG(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0, "adam_multicast: shard bounds must be multiples of 4");
  MMER_CHECK_ARG(((reinterpret_cast<uintptr_t>(p_local) | reinterpret_cast<uintptr_t>(p_mc) | reinterpret_cast<uintptr_t>(g_mc) |
                   reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(shadow_mc) & 7) == 0,
                 "adam_multicast: buffers must be 16-byte aligned");
  if (hi == lo) return 0;
  char buf[32];
  snprintf(buf, sizeof(buf), "%.7g", (double)beta1);
  const double b1 = strtod(buf, nullptr);
  snprintf(buf, sizeof(buf), "%.7g", (double)beta2);
  const double b2 = strtod(buf, nullptr);
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  const long long nt = ((hi - lo) / 4 + ADAM_MC_UNROLL - 1) / ADAM_MC_UNROLL;   // threads: four 16-byte chunks each
  adam_shard_mc_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      p_local, p_mc, g_mc, m, v, (bf16*)shadow_mc, lo, hi, (float)(lr / bc1), beta1, beta2, (float)(1.0 - b1),
      (float)(1.0 - b2), eps, weight_decay, (float)(1.0 / sqrt(bc2)), grad_scale, sumsq_slots, n_slots, max_norm);
  MMER_LAUNCH_CHECK("adam_shard_mc_kernel");
  return 0;
}

int mmer_grad_sumsq(const float* g, int64_t n, float* out, void* stream) {
  MMER_CHECK_ARG(g && out, "grad_sumsq: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "memset(sumsq)");
  if (n <= 0) return 0;
  long long want = (n / 4 + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  sumsq_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(g, n, out);
  MMER_LAUNCH_CHECK("sumsq_kernel");
  return 0;
}

}  // extern "C"
