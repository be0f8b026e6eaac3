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

// ---------------------------------------------------------------------------
// programmatic dependent launch
// ---------------------------------------------------------------------------
// pdl_trigger: the next kernel in the stream (if launched with launch_dep below) may start being scheduled once
// every CTA of this grid has executed it; pdl_wait: block until the previous grid has completed and its memory is
// visible.  Everything a kernel does before pdl_wait (barrier init, TMEM allocation, tensor-map prefetch) overlaps the
// tail of its predecessor; no global memory is touched before it.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch with programmatic stream serialisation (pdl_trigger / pdl_wait above).  ONLY for kernels whose every
// thread executes pdl_wait() before its first global-memory access.  cluster = 2 launches CTA pairs.
extern int g_debug[16];
template <typename... KA, typename... A>
inline cudaError_t launch_dep(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                              A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (!g_debug[MMER_DEBUG_NO_PDL]) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}

// SyncBatchNorm hook (mmer_model.bn_sync / bn_sync_user / bn_world)
struct BnSync {
  int (*fn)(void* user, float* buf, int64_t n, void* stream);
  void* user;
  int world;
};

// cudaFuncSetAttribute is a per-DEVICE setting: a per-kernel bit mask of the devices already configured (the Python layer
// binds a process to one GPU, this keeps a raw C caller on several GPUs correct too)
inline bool needs_func_attr(unsigned long long* done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (*done_mask & bit) return false;
  *done_mask |= bit;
  return true;
}

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
// 8 elements kept as loaded (4 registers for bf16, 8 for fp32): lets a row kernel hold the NEXT row's data in flight
// while it works on the current one without paying fp32 registers for it
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
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
// counter-based dropout.  Elements are grouped in octets (8 consecutive indices, the
// width of one 16-byte bf16 access).  One 32-bit mix of (seed, site, octet index) is
// expanded into four words, i.e. eight 16-bit uniforms; an element is KEPT when its
// uniform >= thr, thr = p * 65536.  Forward and backward regenerate the same decision
// from the element index, so no mask is ever stored.  About 4 integer instructions per
// element, which keeps the GEMM epilogues and the row kernels off the issue limit.
// (torch's Philox stream cannot be matched bit for bit; parity tests run with p = 0 and
// dropout is tested statistically.)  Tensors of up to 2^35 elements per site.
// ---------------------------------------------------------------------------
struct DropCfg {
  uint32_t thr;     // 0 => disabled
  float scale;      // 1 / (1 - thr/65536)
  uint32_t key;     // seed mixed with the site id
  uint32_t thr_hi;  // thr << 16
};
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}
__host__ inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  if (!(p > 0.f)) { d.thr = 0; d.scale = 1.f; d.key = 0; d.thr_hi = 0; return d; }
  uint32_t thr = (uint32_t)(p * 65536.f + 0.5f);
  if (thr > 65535u) thr = 65535u;
  d.thr = thr;
  d.thr_hi = thr << 16;
  d.scale = 65536.f / (float)(65536u - thr);
  d.key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (site + 1u)));
  return d;
}
// well-mixed 32-bit value of one octet
__device__ __forceinline__ uint32_t drop_base(const DropCfg& d, uint32_t octet) {
  uint32_t a = (octet ^ d.key) * 0x9E3779B1u;
  a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15;
  return a;
}
// word k (0..3) of an octet: uniforms of elements 2k (low half) and 2k+1 (high half)
__device__ __forceinline__ uint32_t drop_mult(int k) {
  return k == 0 ? 0x846ca68bU : k == 1 ? 0xc2b2ae35U : k == 2 ? 0x85ebca6bU : 0x27d4eb2fU;
}
__device__ __forceinline__ uint32_t drop_word(uint32_t base, uint32_t mult) {
  uint32_t w = base * mult;
  return w ^ (w >> 16);
}
__device__ __forceinline__ float drop_lo(const DropCfg& d, uint32_t w) { return (w << 16) >= d.thr_hi ? d.scale : 0.f; }
__device__ __forceinline__ float drop_hi(const DropCfg& d, uint32_t w) { return w >= d.thr_hi ? d.scale : 0.f; }
// multiplicative factors for 8 consecutive elements starting at idx (idx % 8 == 0)
__device__ __forceinline__ void drop8(const DropCfg& d, uint64_t idx, float (&f)[8]) {
  const uint32_t a = drop_base(d, (uint32_t)(idx >> 3));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t w = drop_word(a, drop_mult(k));
    f[2 * k] = drop_lo(d, w);
    f[2 * k + 1] = drop_hi(d, w);
  }
}
// factors of the aligned element pair (idx, idx + 1), idx even
__device__ __forceinline__ void drop2(const DropCfg& d, uint64_t idx, float& f0, float& f1) {
  const uint32_t w = drop_word(drop_base(d, (uint32_t)(idx >> 3)), drop_mult((int)((idx >> 1) & 3)));
  f0 = drop_lo(d, w);
  f1 = drop_hi(d, w);
}
// the same factors as drop2(d, octet * 8 + 2 * k) for a caller that knows its octet and hoists mult = drop_mult(k)
// (a lane of an MMA accumulator fragment always owns pair k = lane & 3 of its octets)
__device__ __forceinline__ void drop2_at(const DropCfg& d, uint32_t octet, uint32_t mult, float& f0, float& f1) {
  const uint32_t w = drop_word(drop_base(d, octet), mult);
  f0 = drop_lo(d, w);
  f1 = drop_hi(d, w);
}
__device__ __forceinline__ float drop1(const DropCfg& d, uint64_t idx) {
  const uint32_t w = drop_word(drop_base(d, (uint32_t)(idx >> 3)), drop_mult((int)((idx >> 1) & 3)));
  return (idx & 1) ? drop_hi(d, w) : drop_lo(d, w);
}
// element index of attention probability (head-row `bhi` = (b*H + h)*S + i, key j): rows are padded to a multiple
// of 8 keys so that an octet never straddles two rows (every attention kernel must use this)
__host__ __device__ __forceinline__ int att_drop_stride(int S) { return (S + 7) & ~7; }
__device__ __forceinline__ uint64_t att_drop_index(long long bhi, int j, int stride) {
  return (uint64_t)(bhi * stride + j);
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace mmer
