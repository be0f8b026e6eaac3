// Shared device/host helpers for the sm_100a fusion-step kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmer.h"

namespace mmer {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative code and leaves
// a message for mmer_last_error(); nothing throws, allocates or synchronises.
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
void count_launch(int n = 1);

#define MMER_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      mmer::set_error(__VA_ARGS__);               \
      return MMER_ERR_ARG;                        \
    }                                             \
  } while (0)

#define MMER_LAUNCH_CHECK(what)                                   \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return mmer::cuda_fail(_e, what);      \
    mmer::count_launch();                                         \
  } while (0)

#define MMER_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != 0) return _rc;     \
  } while (0)

// ---------------------------------------------------------------------------
// 8-wide vector access for float / bf16 rows (16 B for bf16, 2 x 16 B for fp32)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// round a value the way it will be stored, so that statistics computed in the
// producing kernel match what a consumer re-reads from memory
template <typename T> __device__ __forceinline__ float round_as(float x) { return to_f(from_f<T>(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// counter-based dropout.  One 32-bit hash of (seed, site, index/2) yields two
// 16-bit uniforms; an element is KEPT when its uniform >= thr, thr = p * 65536.
// Forward and backward regenerate the same decision from the element index, so
// no mask is ever stored.  (torch's Philox stream cannot be matched bit for bit;
// parity tests run with p = 0 and dropout is tested statistically.)
// ---------------------------------------------------------------------------
struct DropCfg {
  uint32_t thr;     // 0 => disabled
  float scale;      // 1 / (1 - thr/65536)
  uint32_t key;     // seed mixed with the site id
};
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}
__host__ inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  if (!(p > 0.f)) { d.thr = 0; d.scale = 1.f; d.key = 0; return d; }
  uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);
  if (thr > 65535u) thr = 65535u;
  d.thr = thr;
  d.scale = 65536.f / (float)(65536u - thr);
  d.key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (site + 1u)));
  return d;
}
// keep-mask bits for the element pair (2*pair, 2*pair+1): bit0, bit1
__device__ __forceinline__ uint32_t drop_pair(const DropCfg& d, uint64_t pair) {
  const uint32_t h = mix32(((uint32_t)pair ^ d.key) + (uint32_t)(pair >> 32) * 0x85EBCA77u);
  return ((h & 0xFFFFu) >= d.thr ? 1u : 0u) | ((h >> 16) >= d.thr ? 2u : 0u);
}
// multiplicative factors for 8 consecutive elements starting at idx (idx % 8 == 0)
__device__ __forceinline__ void drop8(const DropCfg& d, uint64_t idx, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t m = drop_pair(d, (idx >> 1) + i);
    f[2 * i] = (m & 1u) ? d.scale : 0.f;
    f[2 * i + 1] = (m & 2u) ? d.scale : 0.f;
  }
}
__device__ __forceinline__ float drop1(const DropCfg& d, uint64_t idx) {
  uint32_t m = drop_pair(d, idx >> 1);
  return ((m >> (idx & 1)) & 1u) ? d.scale : 0.f;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace mmer
