// Residual + dropout + LayerNorm (+ReLU) forward / backward as bulk-copy-pipelined row kernels.
//
// Replaces, per post-norm encoder sub-layer, `x = norm(x + dropout(sublayer(x)))` of
// nn.TransformerEncoderLayer (configured at train2.py:111-118 / train.py:54-57) and, in the classifier head,
// `Linear -> LayerNorm -> ReLU -> Dropout` (train2.py:217-228), plus their autograd backward.
//
// One persistent CTA per SM: W compute warps + 1 producer warp.  A stage holds W consecutive rows of every input
// array; because consecutive rows are contiguous in HBM, a stage is filled by ONE 1-D bulk (TMA) copy per array
// (W * F elements, 16 KB at W = 16, F = 512, bf16), completion counted on an mbarrier.  Up to four stages are in
// flight, so the memory pipe never waits for the arithmetic.  Warp w owns row w of the stage: a lane reads its
// 16-byte chunks (columns lane*8 + i*256) straight from shared memory, statistics and parameter-gradient partials
// stay in fp32 registers, outputs leave as coalesced 16-byte stores.  Column partials (dgamma, dbeta, the bias
// gradient of the producing Linear) are reduced across the CTA through shared memory and flushed with one fp32
// atomic per column per CTA.
#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

static constexpr float LNP_EPS = 1e-5f;
static constexpr int LNP_MAX_WARPS = 15;   // + 1 producer warp = 512 threads: 128 registers per thread
static constexpr int LNP_MAX_STAGES = 4;

struct LnPipeGeom {
  long long M;
  int F, W, stages, narr;
  uint32_t row_bytes;        // F * sizeof(T)
  uint32_t arr_bytes;        // W * row_bytes      (one array of one stage)
  uint32_t stage_bytes;      // narr * arr_bytes
};

__device__ __forceinline__ float sum8f(const float (&v)[8]) {
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

// producer warp: one elected lane streams the row tiles of this CTA through the stage ring
template <typename T>
__device__ __forceinline__ void lnp_produce(const LnPipeGeom& g, const T* const (&src)[3], uint32_t smem_a, uint32_t full_a,
                                            uint32_t empty_a) {
  const long long tiles = (g.M + g.W - 1) / g.W;
  int stage = 0;
  uint32_t phase = 0;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long row0 = t * g.W;
    const long long left = g.M - row0;
    const uint32_t rows = (uint32_t)(left < g.W ? left : g.W);
    const uint32_t bytes = rows * g.row_bytes;
    mbar_wait(empty_a + 8 * stage, phase ^ 1);
    mbar_expect_tx(full_a + 8 * stage, bytes * g.narr);
    for (int a = 0; a < g.narr; ++a)
      bulk_g2s(smem_a + stage * g.stage_bytes + a * g.arr_bytes, src[a] + row0 * g.F, bytes, full_a + 8 * stage);
    if (++stage == g.stages) { stage = 0; phase ^= 1; }
  }
}

// reduce per-warp column partials across the compute warps and add them to a global vector
template <int NCH>
__device__ __forceinline__ void lnp_flush(float (&part)[NCH][8], float* __restrict__ gout, int F, int W, float* sred,
                                          bool compute_warp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (compute_warp) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        *reinterpret_cast<float4*>(sred + warp * F + c) = make_float4(part[i][0], part[i][1], part[i][2], part[i][3]);
        *reinterpret_cast<float4*>(sred + warp * F + c + 4) = make_float4(part[i][4], part[i][5], part[i][6], part[i][7]);
      }
    }
  }
  __syncthreads();
  if (gout != nullptr) {
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < W; ++w) s += sred[w * F + c];
      atomicAdd(gout + c, s);
    }
  }
}

struct LnSmem {
  uint8_t* data;
  uint32_t data_a, full_a, empty_a;
};
__device__ __forceinline__ LnSmem lnp_setup(const LnPipeGeom& g, uint8_t* smem) {
  LnSmem s;
  s.data = smem;
  s.data_a = smem_u32(smem);
  pdl_trigger();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * g.stage_bytes);
  s.full_a = smem_u32(bars);
  s.empty_a = smem_u32(bars + LNP_MAX_STAGES);
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(s.full_a + 8 * i, 1);
      mbar_init(s.empty_a + 8 * i, g.W);
    }
    mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// forward:  z = x + drop_a(a);  y = drop_y(relu?(LN(z) * gamma + beta));  stats = (mean, rstd)
// ---------------------------------------------------------------------------------------------------------
// MODE < 0: every option decided at run time.  MODE >= 0: a bit set fixed at compile time (LNM_*), which strips the
// predicated code of the unused options from the per-row loop of the hot instances (encoder sub-layers).
enum : int { LNM_X = 1, LNM_RELU = 2, LNM_DROP_A = 4, LNM_DROP_Y = 8, LNM_DBIAS = 16, LNM_ZIN = 32 };

template <typename T, int NCH, int MODE>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
add_ln_fwd_pipe_kernel(const T* __restrict__ a, const T* __restrict__ x, const float* __restrict__ gamma,
                       const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ stats, LnPipeGeom g, int relu_rt,
                       DropCfg da, DropCfg dy) {
  extern __shared__ __align__(128) uint8_t smem[];
  const bool has_x = MODE < 0 ? g.narr > 1 : (MODE & LNM_X) != 0;
  const bool relu = MODE < 0 ? relu_rt != 0 : (MODE & LNM_RELU) != 0;
  const bool drop_a = MODE < 0 ? da.thr != 0 : (MODE & LNM_DROP_A) != 0;
  const bool drop_y = MODE < 0 ? dy.thr != 0 : (MODE & LNM_DROP_Y) != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = lnp_setup(g, smem);
  constexpr bool FULLW = MODE >= 0;   // the specialised instances are launched with F == NCH * 256 only: no column guards
  const int F = FULLW ? NCH * 256 : g.F;
  if (warp == g.W) {
    if (lane == 0) {
      const T* const src[3] = {a, x, nullptr};
      lnp_produce<T>(g, src, sm.data_a, sm.full_a, sm.empty_a);
    }
    return;
  }
  float gm[NCH][8], bt[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
    if (FULLW || c < F) { load8(gamma + c, gm[i]); load8(beta + c, bt[i]); }
  }
  const float invF = 1.f / (float)F;
  const long long tiles = (g.M + g.W - 1) / g.W;
  int stage = 0;
  uint32_t phase = 0;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long row = t * g.W + warp;
    mbar_wait(sm.full_a + 8 * stage, phase);
    float z[NCH][8];
    float s = 0.f;
    if (row < g.M) {
      const T* sa = reinterpret_cast<const T*>(sm.data + (size_t)stage * g.stage_bytes) + warp * F;
      const T* sx = reinterpret_cast<const T*>(sm.data + (size_t)stage * g.stage_bytes + g.arr_bytes) + warp * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (FULLW || c < F) {
          load8(sa + c, z[i]);
          if (drop_a) {
            float f[8];
            drop8(da, (uint64_t)(row * F + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) z[i][j] *= f[j];
          }
          if (has_x) {
            float xv[8];
            load8(sx + c, xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) z[i][j] += xv[j];
          }
          s += sum8f(z[i]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);   // this warp's row is in registers: the slot may be refilled
    if (++stage == g.stages) { stage = 0; phase ^= 1; }
    if (row >= g.M) continue;
    const float mean = warp_sum(s) * invF;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (FULLW || c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invF + LNP_EPS);
    if (lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (FULLW || c < F) {
        const long long off = row * F + c;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = fmaf(fmaf(z[i][j], rstd, nmr), gm[i][j], bt[i][j]);
          if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        if (drop_y) {
          float f[8];
          drop8(dy, (uint64_t)off, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= f[j];
        }
        store8(y + off, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward.  dy: gradient of the output; recomputes z = x + drop_a(a) and xhat from the saved (mean, rstd).
//   dz  = rstd * (g*dy' - mean(g*dy') - xhat * mean(g*dy'*xhat))          (gradient of z: residual path)
//   da  = dz o dropmask_a                                                   (gradient of the sub-layer output)
//   dgamma += sum_rows dy'*xhat, dbeta += sum_rows dy', dbias += sum_rows (stored da)
// ---------------------------------------------------------------------------------------------------------
template <typename T, int NCH, int MODE>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
add_ln_bwd_pipe_kernel(const T* __restrict__ dyp, const T* __restrict__ a, const T* __restrict__ x,
                       const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                       T* __restrict__ dz, T* __restrict__ dap, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dbias, LnPipeGeom g, int relu_rt, DropCfg da, DropCfg dy, int zin_rt) {
  extern __shared__ __align__(128) uint8_t smem[];
  // zin: `a` already holds z = x + dropout(sub-layer output) (written by the fused GEMM + LayerNorm forward, gemm_ln.cu);
  // the dropout factors are then needed only for da = dz o mask
  const bool zin = MODE < 0 ? zin_rt != 0 : (MODE & LNM_ZIN) != 0;
  const bool has_x = MODE < 0 ? g.narr > 2 : (MODE & LNM_X) != 0;
  const bool relu = MODE < 0 ? relu_rt != 0 : (MODE & LNM_RELU) != 0;
  const bool drop_a = MODE < 0 ? da.thr != 0 : (MODE & LNM_DROP_A) != 0;
  const bool drop_y = MODE < 0 ? dy.thr != 0 : (MODE & LNM_DROP_Y) != 0;
  const bool want_dbias = MODE < 0 ? dbias != nullptr : (MODE & LNM_DBIAS) != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = lnp_setup(g, smem);
  constexpr bool FULLW = MODE >= 0;   // the specialised instances are launched with F == NCH * 256 only: no column guards
  const int F = FULLW ? NCH * 256 : g.F;
  const bool compute_warp = warp < g.W;
  float* sgamma = reinterpret_cast<float*>(smem + (size_t)g.stages * g.stage_bytes + 2 * LNP_MAX_STAGES * 8);
  for (int c = threadIdx.x; c < F; c += blockDim.x) sgamma[c] = gamma[c];
  __syncthreads();
  float pg[NCH][8], pb[NCH][8], pbias[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; pbias[i][j] = 0.f; }
  if (!compute_warp) {
    if (lane == 0) {
      const T* const src[3] = {dyp, a, x};
      lnp_produce<T>(g, src, sm.data_a, sm.full_a, sm.empty_a);
    }
  } else {
    const float invF = 1.f / (float)F;
    const long long tiles = (g.M + g.W - 1) / g.W;
    int stage = 0;
    uint32_t phase = 0;
    // (mean, rstd) of a row come from global memory: fetched one tile ahead, so that the ~1 us of DRAM latency is not
    // exposed once per row when the data tile is already waiting in shared memory
    float2 st_next = make_float2(0.f, 0.f);
    if ((long long)blockIdx.x * g.W + warp < g.M)
      st_next = *reinterpret_cast<const float2*>(stats + ((long long)blockIdx.x * g.W + warp) * 2);
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const long long row = t * g.W + warp;
      const float mean = st_next.x, rstd = st_next.y;
      const long long row_next = (t + gridDim.x) * g.W + warp;
      if (row_next < g.M) st_next = *reinterpret_cast<const float2*>(stats + row_next * 2);
      mbar_wait(sm.full_a + 8 * stage, phase);
      float xh[NCH][8], gd[NCH][8], fa[NCH][8];
      float s1 = 0.f, s2 = 0.f;
      if (row < g.M) {
        const uint8_t* sb = sm.data + (size_t)stage * g.stage_bytes;
        const T* sdy = reinterpret_cast<const T*>(sb) + warp * F;
        const T* sa = reinterpret_cast<const T*>(sb + g.arr_bytes) + warp * F;
        const T* sx = reinterpret_cast<const T*>(sb + 2 * g.arr_bytes) + warp * F;
        const float nmr = -mean * rstd;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane * 8 + i * 256;
          if (FULLW || c < F) {
            const long long off = row * F + c;
            float z[8], d[8], gm8[8];
            load8(sgamma + c, gm8);     // gamma lives in shared memory: 16 registers fewer than a per-lane copy
            load8(sa + c, z);
            if (drop_a) {
              drop8(da, (uint64_t)off, fa[i]);
              if (!zin) {
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] *= fa[i][j];
              }
            }
            if (has_x) {
              float xv[8];
              load8(sx + c, xv);
#pragma unroll
              for (int j = 0; j < 8; ++j) z[j] += xv[j];
            }
            load8(sdy + c, d);
            if (drop_y) {
              float f[8];
              drop8(dy, (uint64_t)off, f);
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] *= f[j];
            }
            if (relu) {
              float be[8];
              load8(beta + c, be);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (!(fmaf(fmaf(z[j], rstd, nmr), gm8[j], be[j]) > 0.f)) d[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xh[i][j] = fmaf(z[j], rstd, nmr);
              pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
              pb[i][j] += d[j];
              gd[i][j] = d[j] * gm8[j];
              s1 += gd[i][j];
              s2 = fmaf(gd[i][j], xh[i][j], s2);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);
      if (++stage == g.stages) { stage = 0; phase ^= 1; }
      if (row >= g.M) continue;
      const float c1r = warp_sum(s1) * invF * rstd;
      const float c2r = warp_sum(s2) * invF * rstd;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (FULLW || c < F) {
          const long long off = row * F + c;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[i][j], -c2r, fmaf(gd[i][j], rstd, -c1r));
          store8(dz + off, o);
          if (drop_a) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= fa[i][j];
            if (dap != nullptr) store8(dap + off, o);
          }
          if (want_dbias) {
            // bias gradient of the Linear that produced `a`: column sum of what is stored
#pragma unroll
            for (int j = 0; j < 8; ++j) pbias[i][j] += round_as<T>(o[j]);
          }
        }
      }
    }
  }
  // all stages have been consumed: the ring doubles as the reduction scratch
  float* sred = reinterpret_cast<float*>(smem);
  lnp_flush<NCH>(pg, dgamma, F, g.W, sred, compute_warp);
  lnp_flush<NCH>(pb, dbeta, F, g.W, sred, compute_warp);
  if (want_dbias) lnp_flush<NCH>(pbias, dbias, F, g.W, sred, compute_warp);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int lnp_geometry(long long M, long long F, int elt, int narr, LnPipeGeom* g, size_t* smem_bytes) {
  g->M = M;
  g->F = (int)F;
  g->narr = narr;
  g->row_bytes = (uint32_t)(F * elt);
  const size_t budget = 200 * 1024;
  int W = LNP_MAX_WARPS;
  while (W > 1 && (size_t)2 * W * g->row_bytes * narr > budget) --W;
  if (M < W) W = (int)M;
  g->W = W;
  g->arr_bytes = (uint32_t)W * g->row_bytes;
  g->stage_bytes = g->arr_bytes * narr;
  int stages = (int)(budget / g->stage_bytes);
  if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
  MMER_CHECK_ARG(stages >= 2, "add_ln: row of %lld bytes does not fit the shared-memory pipeline", (long long)g->row_bytes);
  g->stages = stages;
  size_t data = (size_t)stages * g->stage_bytes;
  const size_t red = (size_t)W * F * sizeof(float);   // reduction scratch of the backward kernel
  if (data < red) data = red;
  *smem_bytes = data + 2 * LNP_MAX_STAGES * 8 + (size_t)F * sizeof(float) + 16;   // + gamma copy (backward)
  return 0;
}
static int lnp_grid(const LnPipeGeom& g) {
  const long long tiles = (g.M + g.W - 1) / g.W;
  const long long cap = sm_count();
  return (int)(tiles < cap ? tiles : cap);
}
template <typename K>
static int lnp_set_smem(K kern, size_t smem, size_t* configured) {
  if (smem > *configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(add_ln pipe)");
    *configured = smem;
  }
  return 0;
}

template <typename T, int NCH, int MODE>
static int fwd_launch_mode(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                           const LnPipeGeom& g, size_t smem, int relu, DropCfg da, DropCfg dy, cudaStream_t st) {
  static size_t configured = 0;
  auto kern = add_ln_fwd_pipe_kernel<T, NCH, MODE>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  cudaError_t e = launch_dep(kern, dim3(lnp_grid(g)), dim3((g.W + 1) * 32), smem, st, 1, (const T*)a, (const T*)x, gamma, beta,
                             (T*)y, stats, g, relu, da, dy);
  if (e != cudaSuccess) return cuda_fail(e, "launch(add_ln_fwd_pipe)");
  MMER_LAUNCH_CHECK("add_ln_fwd_pipe_kernel");
  return 0;
}
template <typename T, int NCH>
static int fwd_launch(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                      long long M, long long F, int relu, DropCfg da, DropCfg dy, cudaStream_t st) {
  LnPipeGeom g;
  size_t smem;
  MMER_TRY(lnp_geometry(M, F, sizeof(T), x ? 2 : 1, &g, &smem));
  if (NCH == 2 && sizeof(T) == 2 && F == NCH * 256) {   // the encoder sub-layer instances of the bf16 step
    const int mode = (x ? LNM_X : 0) | (relu ? LNM_RELU : 0) | (da.thr ? LNM_DROP_A : 0) | (dy.thr ? LNM_DROP_Y : 0);
    if (mode == (LNM_X | LNM_DROP_A))
      return fwd_launch_mode<T, NCH, LNM_X | LNM_DROP_A>(x, a, gamma, beta, y, stats, g, smem, relu, da, dy, st);
    if (mode == LNM_X) return fwd_launch_mode<T, NCH, LNM_X>(x, a, gamma, beta, y, stats, g, smem, relu, da, dy, st);
    if (mode == 0)   // LayerNorm of a stored sum z (the GEMM epilogue already added bias, dropout and the residual)
      return fwd_launch_mode<T, NCH, 0>(x, a, gamma, beta, y, stats, g, smem, relu, da, dy, st);
  }
  return fwd_launch_mode<T, NCH, -1>(x, a, gamma, beta, y, stats, g, smem, relu, da, dy, st);
}
template <typename T, int NCH, int MODE>
static int bwd_launch_mode(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                           const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias,
                           const LnPipeGeom& g, size_t smem, int relu, DropCfg da, DropCfg ddy, cudaStream_t st, int zin) {
  static size_t configured = 0;
  auto kern = add_ln_bwd_pipe_kernel<T, NCH, MODE>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  cudaError_t e = launch_dep(kern, dim3(lnp_grid(g)), dim3((g.W + 1) * 32), smem, st, 1, (const T*)dy, (const T*)a, (const T*)x,
                             stats, gamma, beta, (T*)dz, (T*)dap, dgamma, dbeta, dbias, g, relu, da, ddy, zin);
  if (e != cudaSuccess) return cuda_fail(e, "launch(add_ln_bwd_pipe)");
  MMER_LAUNCH_CHECK("add_ln_bwd_pipe_kernel");
  return 0;
}
template <typename T, int NCH>
static int bwd_launch(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                      const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias, long long M,
                      long long F, int relu, DropCfg da, DropCfg ddy, cudaStream_t st, int zin) {
  LnPipeGeom g;
  size_t smem;
  MMER_TRY(lnp_geometry(M, F, sizeof(T), x ? 3 : 2, &g, &smem));
  if (NCH == 2 && sizeof(T) == 2 && F == NCH * 256) {
    const int mode = (x ? LNM_X : 0) | (relu ? LNM_RELU : 0) | (da.thr ? LNM_DROP_A : 0) | (ddy.thr ? LNM_DROP_Y : 0) |
                     (dbias ? LNM_DBIAS : 0) | (zin ? LNM_ZIN : 0);
    if (mode == (LNM_X | LNM_DROP_A | LNM_DBIAS))
      return bwd_launch_mode<T, NCH, LNM_X | LNM_DROP_A | LNM_DBIAS>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta,
                                                                      dbias, g, smem, relu, da, ddy, st, zin);
    if (mode == (LNM_X | LNM_DBIAS))
      return bwd_launch_mode<T, NCH, LNM_X | LNM_DBIAS>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, g, smem,
                                                         relu, da, ddy, st, zin);
    if (mode == (LNM_ZIN | LNM_DROP_A | LNM_DBIAS))   // after the fused GEMM + LayerNorm forward (training step)
      return bwd_launch_mode<T, NCH, LNM_ZIN | LNM_DROP_A | LNM_DBIAS>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta,
                                                                        dbias, g, smem, relu, da, ddy, st, zin);
    if (mode == (LNM_ZIN | LNM_DBIAS))
      return bwd_launch_mode<T, NCH, LNM_ZIN | LNM_DBIAS>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, g,
                                                           smem, relu, da, ddy, st, zin);
  }
  return bwd_launch_mode<T, NCH, -1>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, g, smem, relu, da, ddy, st, zin);
}

#define LNP_DISPATCH(F, CALL)                                    \
  do {                                                           \
    if ((F) <= 256) { constexpr int NCH = 1; return CALL; }      \
    if ((F) <= 512) { constexpr int NCH = 2; return CALL; }      \
    if ((F) <= 1024) { constexpr int NCH = 4; return CALL; }     \
    { constexpr int NCH = 8; return CALL; }                      \
  } while (0)

int add_ln_fwd_pipe(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                    long long M, long long F, int dtype, int relu, DropCfg da, DropCfg dy, cudaStream_t st) {
  if (dtype == MMER_BF16) LNP_DISPATCH(F, (fwd_launch<bf16, NCH>(x, a, gamma, beta, y, stats, M, F, relu, da, dy, st)));
  LNP_DISPATCH(F, (fwd_launch<float, NCH>(x, a, gamma, beta, y, stats, M, F, relu, da, dy, st)));
}
int add_ln_bwd_pipe(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                    const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias, long long M,
                    long long F, int dtype, int relu, DropCfg da, DropCfg ddy, cudaStream_t st, int zin) {
  if (dtype == MMER_BF16)
    LNP_DISPATCH(F, (bwd_launch<bf16, NCH>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, M, F, relu, da, ddy, st, zin)));
  LNP_DISPATCH(F, (bwd_launch<float, NCH>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, M, F, relu, da, ddy, st, zin)));
}

// =========================================================================================================
// Token assembly (train2.py:150-161): x0[b, s] = dropout(LN_v(pv[b, s]) + pos[s])  for s < T,
//                                     x0[b, T] = dropout(LN_a(pa[b])    + pos[T])
// on the same stage ring.  A tile is W consecutive OUTPUT rows r = b*S + s.  Their video sources are a contiguous
// range of pv rows (row r maps to pv row r - b, and the audio row of a sample is exactly the gap between two
// samples), their audio sources a contiguous range of pa rows, so a stage is filled by two bulk copies (three in
// backward, with the gradient rows) regardless of where sample boundaries fall, and every warp has one row per tile.
// =========================================================================================================
// Tile walker: position of a tile's first row as (sample b0, position s0), advanced incrementally so that the
// per-tile bookkeeping needs no 64-bit divisions.
struct EmbedTile {
  long long r0;      // first output row
  int rows;          // output rows in the tile
  int b0, s0;        // sample and position of r0
  long long v0;      // first pv row
  int nv;            // pv rows
  int a0;            // first pa row (= b0: position s0 <= T, so sample b0's audio row is still ahead)
  int na;            // pa rows
};
struct EmbedWalk {
  long long r0, Mtot;
  int b0, s0, W, S, Tn, step_b, step_s;
  __device__ __forceinline__ void init(long long first_tile, long long tiles_step, int W_, long long Mtot_, int S_, int Tn_) {
    W = W_; S = S_; Tn = Tn_; Mtot = Mtot_;
    r0 = first_tile * W;
    b0 = (int)(r0 / S);
    s0 = (int)(r0 % S);
    const long long step = tiles_step * W;
    step_b = (int)(step / S);
    step_s = (int)(step % S);
  }
  __device__ __forceinline__ bool valid() const { return r0 < Mtot; }
  __device__ __forceinline__ void next() {
    r0 += (long long)step_b * S + step_s;
    b0 += step_b;
    s0 += step_s;
    if (s0 >= S) { s0 -= S; ++b0; }
  }
  __device__ __forceinline__ EmbedTile tile() const {
    EmbedTile e;
    e.r0 = r0;
    const long long left = Mtot - r0;
    e.rows = (int)(left < W ? left : W);
    e.b0 = b0; e.s0 = s0;
    int s1 = s0 + e.rows, b1 = b0;
    b1 += s1 / S;
    s1 -= (s1 / S) * S;
    e.v0 = (long long)b0 * Tn + (s0 < Tn ? s0 : Tn);
    e.nv = (int)((long long)b1 * Tn + (s1 < Tn ? s1 : Tn) - e.v0);
    e.a0 = b0;
    e.na = b1 - b0;
    return e;
  }
};
// sample and position of the tile's row w
__device__ __forceinline__ void embed_row(const EmbedTile& e, int w, int S, int& b, int& s) {
  s = e.s0 + w;
  const int q = s / S;
  b = e.b0 + q;
  s -= q * S;
}

struct EmbedGeom {
  long long Mtot;    // B * S output rows
  int F, W, stages, S, Tn, with_grad;
  uint32_t row_bytes, stage_bytes;   // stage = [W rows of sources][W rows of gradients (backward)]
  // position-stable kernels (below): tile stride of a CTA (a multiple of S / gcd(S, W), so that G * W rows are whole
  // samples), first slot of the audio rows and of the gradient rows inside a stage.  G == 0: the general kernels
  // (tile stride = gridDim.x, audio rows packed right behind the video rows).
  int G, aslot, gslot;
};

template <typename T>
__device__ __forceinline__ void embed_produce(const EmbedGeom& g, const T* pv, const T* pa, const T* dx0, uint32_t smem_a,
                                              uint32_t full_a, uint32_t empty_a) {
  int stage = 0;
  uint32_t phase = 0;
  EmbedWalk wk;
  for (wk.init(blockIdx.x, g.G ? g.G : gridDim.x, g.W, g.Mtot, g.S, g.Tn); wk.valid(); wk.next()) {
    const EmbedTile e = wk.tile();
    mbar_wait(empty_a + 8 * stage, phase ^ 1);
    const uint32_t full = full_a + 8 * stage;
    mbar_expect_tx(full, (uint32_t)(e.nv + e.na + (g.with_grad ? e.rows : 0)) * g.row_bytes);
    const uint32_t base = smem_a + stage * g.stage_bytes;
    const uint32_t aslot = g.G ? (uint32_t)g.aslot : (uint32_t)e.nv, gslot = g.G ? (uint32_t)g.gslot : (uint32_t)g.W;
    if (e.nv > 0) bulk_g2s(base, pv + e.v0 * g.F, (uint32_t)e.nv * g.row_bytes, full);
    if (e.na > 0) bulk_g2s(base + aslot * g.row_bytes, pa + (long long)e.a0 * g.F, (uint32_t)e.na * g.row_bytes, full);
    if (g.with_grad) bulk_g2s(base + gslot * g.row_bytes, dx0 + e.r0 * g.F, (uint32_t)e.rows * g.row_bytes, full);
    if (++stage == g.stages) { stage = 0; phase ^= 1; }
  }
}

__device__ __forceinline__ LnSmem embed_setup(const EmbedGeom& g, uint8_t* smem) {
  pdl_trigger();
  LnSmem s;
  s.data = smem;
  s.data_a = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * g.stage_bytes);
  s.full_a = smem_u32(bars);
  s.empty_a = smem_u32(bars + LNP_MAX_STAGES);
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(s.full_a + 8 * i, 1);
      mbar_init(s.empty_a + 8 * i, g.W);
    }
    mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();
  return s;
}

template <typename T, int NCH>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
embed_fwd_pipe_kernel(const T* __restrict__ pv, const T* __restrict__ pa, const float* __restrict__ gv,
                      const float* __restrict__ bv, const float* __restrict__ ga, const float* __restrict__ ba,
                      const float* __restrict__ pos, T* __restrict__ x0, float* __restrict__ stats, EmbedGeom g, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = embed_setup(g, smem);
  const int F = g.F, S = g.S, Tn = g.Tn;
  if (warp == g.W) {
    if (lane == 0) embed_produce<T>(g, pv, pa, nullptr, sm.data_a, sm.full_a, sm.empty_a);
    return;
  }
  float gmv[NCH][8], btv[NCH][8];   // LayerNorm parameters of the video tokens (16 of 17 rows)
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
    if (c < F) { load8(gv + c, gmv[i]); load8(bv + c, btv[i]); }
  }
  const float invF = 1.f / (float)F;
  int stage = 0;
  uint32_t phase = 0;
  EmbedWalk wk;
  for (wk.init(blockIdx.x, gridDim.x, g.W, g.Mtot, S, Tn); wk.valid(); wk.next()) {
    const EmbedTile e = wk.tile();
    const long long row = e.r0 + warp;
    const bool have = warp < e.rows;
    int b, s;
    embed_row(e, warp, S, b, s);
    const bool audio = s == Tn;
    mbar_wait(sm.full_a + 8 * stage, phase);
    float z[NCH][8];
    float sum = 0.f;
    if (have) {
      const int slot = audio ? e.nv + (b - e.a0) : (int)((long long)b * Tn + s - e.v0);
      const T* src = reinterpret_cast<const T*>(sm.data + (size_t)stage * g.stage_bytes) + (size_t)slot * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) { load8(src + c, z[i]); sum += sum8f(z[i]); }
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);
    if (++stage == g.stages) { stage = 0; phase ^= 1; }
    if (!have) continue;
    const float mean = warp_sum(sum) * invF;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invF + LNP_EPS);
    if (lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        float pe[8], o[8];
        load8(pos + (long long)s * F + c, pe);
        if (audio) {     // warp-uniform: 1 row in S
          float gg[8], bb[8];
          load8(ga + c, gg);
          load8(ba + c, bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(z[i][j], rstd, nmr), gg[j], bb[j]) + pe[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(z[i][j], rstd, nmr), gmv[i][j], btv[i][j]) + pe[j];
        }
        if (dc.thr) {
          float f[8];
          drop8(dc, (uint64_t)off, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= f[j];
        }
        store8(x0 + off, o);
      }
    }
  }
}

// backward of the token assembly: dpv / dpa = LayerNorm backward of the (dropout-masked) gradient rows, dgamma of both
// norms.  dbeta of both norms and dpos are column sums of the masked gradient per position: embed_dpos_kernel.
template <typename T, int NCH>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
embed_bwd_pipe_kernel(const T* __restrict__ dx0, const T* __restrict__ pv, const T* __restrict__ pa,
                      const float* __restrict__ stats, const float* __restrict__ gv, const float* __restrict__ ga,
                      T* __restrict__ dpv, T* __restrict__ dpa, float* __restrict__ dgv, float* __restrict__ dga,
                      float* __restrict__ dbias_v, float* __restrict__ dbias_a, EmbedGeom g, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = embed_setup(g, smem);
  const int F = g.F, S = g.S, Tn = g.Tn;
  const bool compute_warp = warp < g.W;
  // audio rows are 1 in S: their dgamma partials go through shared-memory atomics instead of a second register set
  float* sga = reinterpret_cast<float*>(smem + (size_t)g.stages * g.stage_bytes + 2 * LNP_MAX_STAGES * 8);
  float* sba = sga + F;     // audio rows' contribution to the audio projection's bias gradient
  for (int c = threadIdx.x; c < 2 * F; c += blockDim.x) sga[c] = 0.f;
  __syncthreads();
  float pg[NCH][8], pbv[NCH][8];   // dgamma and projection-bias-gradient partials of the video rows
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pbv[i][j] = 0.f; }
  if (!compute_warp) {
    if (lane == 0) embed_produce<T>(g, pv, pa, dx0, sm.data_a, sm.full_a, sm.empty_a);
  } else {
    float gmv[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) load8(gv + c, gmv[i]);
    }
    const float invF = 1.f / (float)F;
    int stage = 0;
    uint32_t phase = 0;
    EmbedWalk wk;
    for (wk.init(blockIdx.x, gridDim.x, g.W, g.Mtot, S, Tn); wk.valid(); wk.next()) {
      const EmbedTile e = wk.tile();
      const long long row = e.r0 + warp;
      const bool have = warp < e.rows;
      int b, s;
      embed_row(e, warp, S, b, s);
      const bool audio = s == Tn;
      float mean = 0.f, rstd = 0.f;
      if (have) {
        const float2 st = *reinterpret_cast<const float2*>(stats + row * 2);
        mean = st.x;
        rstd = st.y;
      }
      mbar_wait(sm.full_a + 8 * stage, phase);
      float xh[NCH][8], gd[NCH][8];
      float s1 = 0.f, s2 = 0.f;
      if (have) {
        const uint8_t* sb = sm.data + (size_t)stage * g.stage_bytes;
        const int slot = audio ? e.nv + (b - e.a0) : (int)((long long)b * Tn + s - e.v0);
        const T* sz = reinterpret_cast<const T*>(sb) + (size_t)slot * F;
        const T* sd = reinterpret_cast<const T*>(sb + (size_t)g.W * g.row_bytes) + (size_t)warp * F;
        const float nmr = -mean * rstd;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane * 8 + i * 256;
          if (c < F) {
            float d[8], z[8], gg[8];
            load8(sd + c, d);
            if (dc.thr) {
              float f[8];
              drop8(dc, (uint64_t)(row * F + c), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] *= f[j];
            }
            load8(sz + c, z);
            if (audio) {     // warp-uniform: 1 row in S
              load8(ga + c, gg);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                xh[i][j] = fmaf(z[j], rstd, nmr);
                atomicAdd(sga + c + j, d[j] * xh[i][j]);
                gd[i][j] = d[j] * gg[j];
                s1 += gd[i][j];
                s2 = fmaf(gd[i][j], xh[i][j], s2);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                xh[i][j] = fmaf(z[j], rstd, nmr);
                pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
                gd[i][j] = d[j] * gmv[i][j];
                s1 += gd[i][j];
                s2 = fmaf(gd[i][j], xh[i][j], s2);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);
      if (++stage == g.stages) { stage = 0; phase ^= 1; }
      if (!have) continue;
      const float c1r = warp_sum(s1) * invF * rstd;
      const float c2r = warp_sum(s2) * invF * rstd;
      T* dst = audio ? dpa + (long long)b * F : dpv + ((long long)b * Tn + s) * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[i][j], -c2r, fmaf(gd[i][j], rstd, -c1r));
          store8(dst + c, o);
          if (audio) {
            if (dbias_a != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) atomicAdd(sba + c + j, round_as<T>(o[j]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) pbv[i][j] += round_as<T>(o[j]);
          }
        }
      }
    }
  }
  float* sred = reinterpret_cast<float*>(smem);
  lnp_flush<NCH>(pg, dgv, F, g.W, sred, compute_warp);
  if (dbias_v != nullptr) lnp_flush<NCH>(pbv, dbias_v, F, g.W, sred, compute_warp);
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    atomicAdd(dga + c, sga[c]);
    if (dbias_a != nullptr) atomicAdd(dbias_a + c, sba[c]);
  }
}

// dpos[s][c] += sum_b d[b,s,c];  dbeta_video[c] += the same for s < T;  dbeta_audio[c] += for s == T, with d the
// dropout-masked gradient of the assembled tokens: a column sum of dx0 viewed as [B][S*F].
template <typename T>
__global__ void __launch_bounds__(256)
embed_dpos_kernel(const T* __restrict__ dx0, float* __restrict__ dpos, float* __restrict__ dbv, float* __restrict__ dba,
                  int B, int S, int F, DropCfg dc) {
  __shared__ float sred[8][256];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long N = (long long)S * F;
  const long long c = (long long)blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    for (int b = blockIdx.y * 8 + warp; b < B; b += gridDim.y * 8) {
      float v[8];
      const long long off = (long long)b * N + c;
      load8(dx0 + off, v);
      if (dc.thr) {
        float f[8];
        drop8(dc, (uint64_t)off, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sred[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const long long cc = (long long)blockIdx.x * 256 + threadIdx.x;
  if (cc < N) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sacc += sred[w][threadIdx.x];
    if (dpos != nullptr) atomicAdd(dpos + cc, sacc);
    const int s = (int)(cc / F), col = (int)(cc % F);
    float* db = s == S - 1 ? dba : dbv;
    if (db != nullptr) atomicAdd(db + col, sacc);
  }
}

// =========================================================================================================
// Position-stable token assembly.  When the tile stride of a CTA, G tiles of W rows, is a whole number of samples
// (G * W = k * S), row w of EVERY tile a CTA sees sits at the same position s of its sample, k samples further on.
// A warp is then bound to one position for the whole kernel: pos_embed[s] and the LayerNorm parameters of its
// modality (video / audio) live in registers, the slot of its source row inside a stage is a constant, nothing is
// divided or fetched from global memory inside the loop, and in backward the column sums per position (dpos, both
// dbeta) are register partials flushed once -- the separate embed_dpos pass over the gradient (a third of the
// backward traffic) disappears.  The price is a grid of G <= SMs CTAs (136 of 148 at S = 17, W = 15).
// Stage layout: [W video slots][NA audio slots][W gradient rows (backward)], NA = W / S + 2.
// =========================================================================================================
struct EmbedPosWarp {
  long long row0;   // first output row of this warp
  long long step;   // rows between two tiles of the CTA
  int b0, kb;       // its sample in the first tile, samples per step
  int s, slot;      // position, source slot inside a stage
  bool audio;
  __device__ __forceinline__ void init(const EmbedGeom& g, int warp) {
    const long long r00 = (long long)blockIdx.x * g.W;
    const int b00 = (int)(r00 / g.S), s00 = (int)(r00 % g.S);
    const int q = (s00 + warp) / g.S;
    s = s00 + warp - q * g.S;
    audio = s == g.Tn;
    slot = audio ? g.aslot + q : q * g.Tn + s - (s00 < g.Tn ? s00 : g.Tn);
    row0 = r00 + warp;
    step = (long long)g.G * g.W;
    b0 = b00 + q;
    kb = (int)(step / g.S);
  }
};

// FULL: F == NCH * 256 (no column guards); DROP: dropout enabled (dc.thr != 0).  The dropout scale is folded into the
// warp's constants: out = keep ? xhat * (gamma * scale) + (beta + pos) * scale : 0, one select per element.
template <typename T, int NCH, bool FULL, bool DROP>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
embed_fwd_pos_kernel(const T* __restrict__ pv, const T* __restrict__ pa, const float* __restrict__ gv,
                     const float* __restrict__ bv, const float* __restrict__ ga, const float* __restrict__ ba,
                     const float* __restrict__ pos, T* __restrict__ x0, float* __restrict__ stats, EmbedGeom g, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = embed_setup(g, smem);
  const int F = FULL ? NCH * 256 : g.F;
  if (warp == g.W) {
    if (lane == 0) embed_produce<T>(g, pv, pa, nullptr, sm.data_a, sm.full_a, sm.empty_a);
    return;
  }
  EmbedPosWarp me;
  me.init(g, warp);
  float gm[NCH][8], bp[NCH][8];   // this warp's gamma * scale and (beta + pos_embed[s]) * scale
  {
    const float* gsrc = me.audio ? ga : gv;
    const float* bsrc = me.audio ? ba : bv;
    const float sc = DROP ? dc.scale : 1.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (FULL || c < F) {
        float pe[8];
        load8(gsrc + c, gm[i]);
        load8(bsrc + c, bp[i]);
        load8(pos + (long long)me.s * F + c, pe);
#pragma unroll
        for (int j = 0; j < 8; ++j) { gm[i][j] *= sc; bp[i][j] = (bp[i][j] + pe[j]) * sc; }
      }
    }
  }
  const float invF = 1.f / (float)F;
  int stage = 0;
  uint32_t phase = 0;
  long long row = me.row0;
  for (long long r0 = (long long)blockIdx.x * g.W; r0 < g.Mtot; r0 += me.step, row += me.step) {
    const bool have = row < g.Mtot;
    mbar_wait(sm.full_a + 8 * stage, phase);
    float z[NCH][8];
    float sum = 0.f;
    if (have) {
      const T* src = reinterpret_cast<const T*>(sm.data + (size_t)stage * g.stage_bytes) + (size_t)me.slot * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (FULL || c < F) { load8(src + c, z[i]); sum += sum8f(z[i]); }
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);
    if (++stage == g.stages) { stage = 0; phase ^= 1; }
    if (!have) continue;
    const float mean = warp_sum(sum) * invF;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (FULL || c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invF + LNP_EPS);
    if (lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (FULL || c < F) {
        const long long off = row * F + c;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(z[i][j], rstd, nmr), gm[i][j], bp[i][j]);
        if (DROP) {
          const uint32_t a = drop_base(dc, (uint32_t)((uint64_t)off >> 3));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t w = drop_word(a, drop_mult(k));
            if (!((w << 16) >= dc.thr_hi)) o[2 * k] = 0.f;
            if (!(w >= dc.thr_hi)) o[2 * k + 1] = 0.f;
          }
        }
        store8(x0 + off, o);
      }
    }
  }
}

// Column partials of the compute warps -> global vectors, video and audio warps apart; optionally every warp's own
// partial into row `wpos[w]` of a [S][F] matrix (dpos).  `sred` is [W][F] fp32 scratch, `waud` / `wpos` per-warp flags.
template <int NCH>
__device__ __forceinline__ void embed_pos_flush(float (&part)[NCH][8], float* __restrict__ gvid, float* __restrict__ gaud,
                                                float* __restrict__ gpos, int F, int W, float* sred, const int* waud,
                                                const int* wpos, bool compute_warp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (compute_warp) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        *reinterpret_cast<float4*>(sred + warp * F + c) = make_float4(part[i][0], part[i][1], part[i][2], part[i][3]);
        *reinterpret_cast<float4*>(sred + warp * F + c + 4) = make_float4(part[i][4], part[i][5], part[i][6], part[i][7]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float sv = 0.f, sa = 0.f;
    bool any_a = false;
    for (int w = 0; w < W; ++w) {
      const float v = sred[w * F + c];
      if (waud[w]) { sa += v; any_a = true; } else sv += v;
    }
    if (gvid != nullptr) atomicAdd(gvid + c, sv);
    if (gaud != nullptr && any_a) atomicAdd(gaud + c, sa);
  }
  if (gpos != nullptr) {
    const int F4 = F >> 2;
    for (int idx = threadIdx.x; idx < W * F4; idx += blockDim.x) {
      const int w = idx / F4, c = (idx - w * F4) * 4;
      if (wpos[w] >= 0)
        atomicAdd(reinterpret_cast<float4*>(gpos + (long long)wpos[w] * F + c), *reinterpret_cast<const float4*>(sred + w * F + c));
    }
  }
}

template <typename T, int NCH, bool FULL, bool DROP>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
embed_bwd_pos_kernel(const T* __restrict__ dx0, const T* __restrict__ pv, const T* __restrict__ pa,
                     const float* __restrict__ stats, const float* __restrict__ gv, const float* __restrict__ ga,
                     T* __restrict__ dpv, T* __restrict__ dpa, float* __restrict__ dgv, float* __restrict__ dbv,
                     float* __restrict__ dga, float* __restrict__ dba, float* __restrict__ dpos,
                     float* __restrict__ dbias_v, float* __restrict__ dbias_a, EmbedGeom g, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = embed_setup(g, smem);
  const int F = FULL ? NCH * 256 : g.F, Tn = g.Tn;
  const bool compute_warp = warp < g.W;
  // both gamma vectors in shared memory (a per-lane register copy would cost 16 registers), per-warp flags behind them
  float* sgam = reinterpret_cast<float*>(smem + (size_t)g.stages * g.stage_bytes + 2 * LNP_MAX_STAGES * 8);
  int* waud = reinterpret_cast<int*>(sgam + 2 * F);
  int* wpos = waud + LNP_MAX_WARPS + 1;
  for (int c = threadIdx.x; c < F; c += blockDim.x) { sgam[c] = gv[c]; sgam[F + c] = ga[c]; }
  EmbedPosWarp me;
  me.init(g, compute_warp ? warp : 0);
  if (compute_warp && lane == 0) {
    waud[warp] = me.audio ? 1 : 0;
    wpos[warp] = me.row0 < g.Mtot ? me.s : -1;   // a warp without a single row has nothing to add to dpos
  }
  __syncthreads();
  float pg[NCH][8], pb[NCH][8], pbias[NCH][8];   // dgamma, dbeta (= dpos of this position), projection bias gradient
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; pbias[i][j] = 0.f; }
  if (!compute_warp) {
    if (lane == 0) embed_produce<T>(g, pv, pa, dx0, sm.data_a, sm.full_a, sm.empty_a);
  } else {
    const float* gam = sgam + (me.audio ? F : 0);
    const bool want_bias = (me.audio ? dbias_a : dbias_v) != nullptr;
    const float invF = 1.f / (float)F;
    int stage = 0;
    uint32_t phase = 0;
    long long row = me.row0;
    int b = me.b0;
    float2 st_next = make_float2(0.f, 0.f);
    if (row < g.Mtot) st_next = *reinterpret_cast<const float2*>(stats + row * 2);
    for (long long r0 = (long long)blockIdx.x * g.W; r0 < g.Mtot; r0 += me.step, row += me.step, b += me.kb) {
      const bool have = row < g.Mtot;
      const float mean = st_next.x, rstd = st_next.y;
      if (row + me.step < g.Mtot) st_next = *reinterpret_cast<const float2*>(stats + (row + me.step) * 2);
      mbar_wait(sm.full_a + 8 * stage, phase);
      float xh[NCH][8], gd[NCH][8];
      float s1 = 0.f, s2 = 0.f;
      if (have) {
        const uint8_t* sb = sm.data + (size_t)stage * g.stage_bytes;
        const T* sz = reinterpret_cast<const T*>(sb) + (size_t)me.slot * F;
        const T* sd = reinterpret_cast<const T*>(sb) + (size_t)(g.gslot + warp) * F;
        const float nmr = -mean * rstd;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane * 8 + i * 256;
          if (FULL || c < F) {
            float d[8], z[8], gg[8];
            load8(sd + c, d);
            if (DROP) {
              float f[8];
              drop8(dc, (uint64_t)(row * F + c), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] *= f[j];
            }
            load8(sz + c, z);
            load8(gam + c, gg);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xh[i][j] = fmaf(z[j], rstd, nmr);
              pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
              pb[i][j] += d[j];
              gd[i][j] = d[j] * gg[j];
              s1 += gd[i][j];
              s2 = fmaf(gd[i][j], xh[i][j], s2);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty_a + 8 * stage);
      if (++stage == g.stages) { stage = 0; phase ^= 1; }
      if (!have) continue;
      const float c1r = warp_sum(s1) * invF * rstd;
      const float c2r = warp_sum(s2) * invF * rstd;
      T* dst = me.audio ? dpa + (long long)b * F : dpv + ((long long)b * Tn + me.s) * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (FULL || c < F) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[i][j], -c2r, fmaf(gd[i][j], rstd, -c1r));
          store8(dst + c, o);
          if (want_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) pbias[i][j] += round_as<T>(o[j]);
          }
        }
      }
    }
  }
  float* sred = reinterpret_cast<float*>(smem);   // every stage has been consumed: the ring is the reduction scratch
  embed_pos_flush<NCH>(pg, dgv, dga, nullptr, F, g.W, sred, waud, wpos, compute_warp);
  embed_pos_flush<NCH>(pb, dbv, dba, dpos, F, g.W, sred, waud, wpos, compute_warp);
  if (dbias_v != nullptr || dbias_a != nullptr)
    embed_pos_flush<NCH>(pbias, dbias_v, dbias_a, nullptr, F, g.W, sred, waud, wpos, compute_warp);
}

static int embed_geometry(long long B, long long T, long long F, int elt, int with_grad, EmbedGeom* g, size_t* smem_bytes) {
  g->Mtot = B * (T + 1);
  g->F = (int)F; g->S = (int)T + 1; g->Tn = (int)T; g->with_grad = with_grad;
  g->G = 0; g->aslot = 0; g->gslot = 0;
  g->row_bytes = (uint32_t)(F * elt);
  const size_t budget = 200 * 1024;
  const int arrays = with_grad ? 2 : 1;
  int W = LNP_MAX_WARPS;
  while (W > 1 && (size_t)2 * W * g->row_bytes * arrays > budget) --W;
  if (g->Mtot < W) W = (int)g->Mtot;
  g->W = W;
  g->stage_bytes = (uint32_t)W * g->row_bytes * arrays;
  int stages = (int)(budget / g->stage_bytes);
  if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
  MMER_CHECK_ARG(stages >= 2, "embed: row of %lld bytes does not fit the shared-memory pipeline", (long long)g->row_bytes);
  g->stages = stages;
  size_t data = (size_t)stages * g->stage_bytes;
  const size_t red = (size_t)W * F * sizeof(float);
  if (data < red) data = red;
  *smem_bytes = data + 2 * LNP_MAX_STAGES * 8 + (size_t)2 * F * sizeof(float) + 16;
  return 0;
}

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// Plan of the position-stable kernels: the warp count W (<= 15) and tile stride G (<= SMs, a multiple of
// S / gcd(S, W)) with the most rows in flight.  Returns false when no plan keeps at least 85 % of the rows the
// general kernel has in flight (long sequences: S / gcd(S, W) exceeds the SM count), for rows wider than 512 columns,
// or when the debug knob asks for the general kernels.
static bool embed_pos_plan(long long B, long long T, long long F, int elt, int with_grad, EmbedGeom* g, size_t* smem_bytes) {
  if (g_debug[MMER_DEBUG_EMBED_GENERIC]) return false;
  if (F > 512) return false;   // wider rows: the per-warp parameter registers of these kernels would spill
  const int S = (int)T + 1;
  const long long Mtot = B * S;
  const uint32_t row_bytes = (uint32_t)(F * elt);
  const size_t budget = 200 * 1024;
  const int cap = sm_count();
  int bestW = 0, bestG = 0, maxW = 0;
  for (int W = LNP_MAX_WARPS; W >= 1; --W) {
    const size_t stage = (size_t)(W + W / S + 2 + (with_grad ? W : 0)) * row_bytes;
    if (2 * stage > budget) continue;
    if (maxW == 0) maxW = W;
    const int g0 = S / gcd_int(S, W);
    if (g0 > cap) continue;
    const int G = (cap / g0) * g0;
    if ((long long)G * W > (long long)bestG * bestW) { bestG = G; bestW = W; }
  }
  if (bestW == 0 || (long long)bestG * bestW * 100 < (long long)cap * maxW * 85) return false;
  g->Mtot = Mtot;
  g->F = (int)F; g->S = S; g->Tn = (int)T; g->with_grad = with_grad;
  g->row_bytes = row_bytes;
  g->W = bestW;
  g->G = bestG;
  g->aslot = bestW;
  g->gslot = bestW + bestW / S + 2;
  g->stage_bytes = (uint32_t)(g->gslot + (with_grad ? bestW : 0)) * row_bytes;
  int stages = (int)(budget / g->stage_bytes);
  if (stages > LNP_MAX_STAGES) stages = LNP_MAX_STAGES;
  g->stages = stages;
  size_t data = (size_t)stages * g->stage_bytes;
  const size_t red = (size_t)bestW * F * sizeof(float);
  if (data < red) data = red;
  *smem_bytes = data + 2 * LNP_MAX_STAGES * 8 + (size_t)2 * F * sizeof(float) + 2 * (LNP_MAX_WARPS + 1) * sizeof(int) + 16;
  return true;
}

template <typename T, int NCH>
static int embed_fwd_launch(const void* pv, const void* pa, const float* gv, const float* bv, const float* ga,
                            const float* ba, const float* pos, void* x0, float* stats, long long B, long long T_, long long F,
                            DropCfg dc, cudaStream_t st) {
  EmbedGeom g;
  size_t smem;
  if constexpr (NCH <= 2) {
    if (embed_pos_plan(B, T_, F, sizeof(T), 0, &g, &smem)) {
      const bool full = F == NCH * 256, drop = dc.thr != 0;
      auto kpos = full ? (drop ? embed_fwd_pos_kernel<T, NCH, true, true> : embed_fwd_pos_kernel<T, NCH, true, false>)
                       : (drop ? embed_fwd_pos_kernel<T, NCH, false, true> : embed_fwd_pos_kernel<T, NCH, false, false>);
      static size_t configured_pos[4] = {0, 0, 0, 0};
      MMER_TRY(lnp_set_smem(kpos, smem, &configured_pos[(full ? 2 : 0) + (drop ? 1 : 0)]));
      const long long tiles = (g.Mtot + g.W - 1) / g.W;
      cudaError_t e = launch_dep(kpos, dim3((unsigned)(tiles < g.G ? tiles : g.G)), dim3((g.W + 1) * 32), smem, st, 1,
                                 (const T*)pv, (const T*)pa, gv, bv, ga, ba, pos, (T*)x0, stats, g, dc);
      if (e != cudaSuccess) return cuda_fail(e, "launch(embed_fwd_pos)");
      MMER_LAUNCH_CHECK("embed_fwd_pos_kernel");
      return 0;
    }
  }
  MMER_TRY(embed_geometry(B, T_, F, sizeof(T), 0, &g, &smem));
  static size_t configured = 0;
  auto kern = embed_fwd_pipe_kernel<T, NCH>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  const long long tiles = (g.Mtot + g.W - 1) / g.W;
  const long long cap = sm_count();
  cudaError_t e = launch_dep(kern, dim3((unsigned)(tiles < cap ? tiles : cap)), dim3((g.W + 1) * 32), smem, st, 1, (const T*)pv,
                             (const T*)pa, gv, bv, ga, ba, pos, (T*)x0, stats, g, dc);
  if (e != cudaSuccess) return cuda_fail(e, "launch(embed_fwd_pipe)");
  MMER_LAUNCH_CHECK("embed_fwd_pipe_kernel");
  return 0;
}
template <typename T, int NCH>
static int embed_bwd_launch(const void* dx0, const void* pv, const void* pa, const float* stats, const float* gv,
                            const float* ga, void* dpv, void* dpa, float* dgv, float* dbv, float* dga, float* dba,
                            float* dpos, float* dbias_v, float* dbias_a, long long B, long long T_, long long F, DropCfg dc,
                            cudaStream_t st) {
  EmbedGeom g;
  size_t smem;
  if constexpr (NCH <= 2) {
    if (embed_pos_plan(B, T_, F, sizeof(T), 1, &g, &smem)) {
      const bool full = F == NCH * 256, drop = dc.thr != 0;
      auto kpos = full ? (drop ? embed_bwd_pos_kernel<T, NCH, true, true> : embed_bwd_pos_kernel<T, NCH, true, false>)
                       : (drop ? embed_bwd_pos_kernel<T, NCH, false, true> : embed_bwd_pos_kernel<T, NCH, false, false>);
      static size_t configured_pos[4] = {0, 0, 0, 0};
      MMER_TRY(lnp_set_smem(kpos, smem, &configured_pos[(full ? 2 : 0) + (drop ? 1 : 0)]));
      const long long tiles = (g.Mtot + g.W - 1) / g.W;
      cudaError_t e = launch_dep(kpos, dim3((unsigned)(tiles < g.G ? tiles : g.G)), dim3((g.W + 1) * 32), smem, st, 1,
                                 (const T*)dx0, (const T*)pv, (const T*)pa, stats, gv, ga, (T*)dpv, (T*)dpa, dgv, dbv, dga, dba,
                                 dpos, dbias_v, dbias_a, g, dc);
      if (e != cudaSuccess) return cuda_fail(e, "launch(embed_bwd_pos)");
      MMER_LAUNCH_CHECK("embed_bwd_pos_kernel");
      return 0;
    }
  }
  MMER_TRY(embed_geometry(B, T_, F, sizeof(T), 1, &g, &smem));
  static size_t configured = 0;
  auto kern = embed_bwd_pipe_kernel<T, NCH>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  const long long tiles = (g.Mtot + g.W - 1) / g.W;
  const long long cap = sm_count();
  cudaError_t e = launch_dep(kern, dim3((unsigned)(tiles < cap ? tiles : cap)), dim3((g.W + 1) * 32), smem, st, 1, (const T*)dx0,
                             (const T*)pv, (const T*)pa, stats, gv, ga, (T*)dpv, (T*)dpa, dgv, dga, dbias_v, dbias_a, g, dc);
  if (e != cudaSuccess) return cuda_fail(e, "launch(embed_bwd_pipe)");
  MMER_LAUNCH_CHECK("embed_bwd_pipe_kernel");
  const long long N = (T_ + 1) * F;
  int gy = (int)((sm_count() * 4 + (N + 255) / 256 - 1) / ((N + 255) / 256));
  if (gy < 1) gy = 1;
  if (gy > (B + 7) / 8) gy = (int)((B + 7) / 8);
  e = launch_dep(embed_dpos_kernel<T>, dim3((unsigned)((N + 255) / 256), (unsigned)gy), dim3(256), 0, st, 1, (const T*)dx0, dpos,
                 dbv, dba, (int)B, (int)T_ + 1, (int)F, dc);
  if (e != cudaSuccess) return cuda_fail(e, "launch(embed_dpos)");
  MMER_LAUNCH_CHECK("embed_dpos_kernel");
  return 0;
}

// LayerNorm variant of the token assembly (train2.py); the BatchNorm variant (no norm here) stays in rowops.cu
int embed_fwd_pipe(const void* pv, const void* pa, const float* gv, const float* bv, const float* ga, const float* ba,
                   const float* pos, void* x0, float* stats, long long B, long long T, long long F, int dtype, DropCfg dc,
                   cudaStream_t st) {
  if (dtype == MMER_BF16) LNP_DISPATCH(F, (embed_fwd_launch<bf16, NCH>(pv, pa, gv, bv, ga, ba, pos, x0, stats, B, T, F, dc, st)));
  LNP_DISPATCH(F, (embed_fwd_launch<float, NCH>(pv, pa, gv, bv, ga, ba, pos, x0, stats, B, T, F, dc, st)));
}
int embed_bwd_pipe(const void* dx0, const void* pv, const void* pa, const float* stats, const float* gv, const float* ga,
                   void* dpv, void* dpa, float* dgv, float* dbv, float* dga, float* dba, float* dpos, float* dbias_v,
                   float* dbias_a, long long B, long long T, long long F, int dtype, DropCfg dc, cudaStream_t st) {
  if (dtype == MMER_BF16)
    LNP_DISPATCH(F, (embed_bwd_launch<bf16, NCH>(dx0, pv, pa, stats, gv, ga, dpv, dpa, dgv, dbv, dga, dba, dpos, dbias_v,
                                                 dbias_a, B, T, F, dc, st)));
  LNP_DISPATCH(F, (embed_bwd_launch<float, NCH>(dx0, pv, pa, stats, gv, ga, dpv, dpa, dgv, dbv, dga, dba, dpos, dbias_v,
                                                dbias_a, B, T, F, dc, st)));
}

}  // namespace mmer
