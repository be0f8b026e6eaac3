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
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// forward:  z = x + drop_a(a);  y = drop_y(relu?(LN(z) * gamma + beta));  stats = (mean, rstd)
// ---------------------------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
add_ln_fwd_pipe_kernel(const T* __restrict__ a, const T* __restrict__ x, const float* __restrict__ gamma,
                       const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ stats, LnPipeGeom g, int relu,
                       DropCfg da, DropCfg dy) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = lnp_setup(g, smem);
  const int F = g.F;
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
    if (c < F) { load8(gamma + c, gm[i]); load8(beta + c, bt[i]); }
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
        if (c < F) {
          load8(sa + c, z[i]);
          if (da.thr) {
            float f[8];
            drop8(da, (uint64_t)(row * F + c), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) z[i][j] *= f[j];
          }
          if (g.narr > 1) {
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
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = fmaf(fmaf(z[i][j], rstd, nmr), gm[i][j], bt[i][j]);
          if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        if (dy.thr) {
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
template <typename T, int NCH>
__global__ void __launch_bounds__((LNP_MAX_WARPS + 1) * 32, 1)
add_ln_bwd_pipe_kernel(const T* __restrict__ dyp, const T* __restrict__ a, const T* __restrict__ x,
                       const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                       T* __restrict__ dz, T* __restrict__ dap, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dbias, LnPipeGeom g, int relu, DropCfg da, DropCfg dy) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LnSmem sm = lnp_setup(g, smem);
  const int F = g.F;
  const bool compute_warp = warp < g.W;
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
    float gm[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) load8(gamma + c, gm[i]);
    }
    const float invF = 1.f / (float)F;
    const long long tiles = (g.M + g.W - 1) / g.W;
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const long long row = t * g.W + warp;
      float mean = 0.f, rstd = 0.f;
      if (row < g.M) {
        const float2 st = *reinterpret_cast<const float2*>(stats + row * 2);
        mean = st.x;
        rstd = st.y;
      }
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
          if (c < F) {
            const long long off = row * F + c;
            float z[8], d[8];
            load8(sa + c, z);
            if (da.thr) {
              drop8(da, (uint64_t)off, fa[i]);
#pragma unroll
              for (int j = 0; j < 8; ++j) z[j] *= fa[i][j];
            }
            if (g.narr > 2) {
              float xv[8];
              load8(sx + c, xv);
#pragma unroll
              for (int j = 0; j < 8; ++j) z[j] += xv[j];
            }
            load8(sdy + c, d);
            if (dy.thr) {
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
                if (!(fmaf(fmaf(z[j], rstd, nmr), gm[i][j], be[j]) > 0.f)) d[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xh[i][j] = fmaf(z[j], rstd, nmr);
              pg[i][j] = fmaf(d[j], xh[i][j], pg[i][j]);
              pb[i][j] += d[j];
              gd[i][j] = d[j] * gm[i][j];
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
        if (c < F) {
          const long long off = row * F + c;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[i][j], -c2r, fmaf(gd[i][j], rstd, -c1r));
          store8(dz + off, o);
          if (da.thr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= fa[i][j];
            if (dap != nullptr) store8(dap + off, o);
          }
          if (dbias != nullptr) {
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
  if (dbias != nullptr) lnp_flush<NCH>(pbias, dbias, F, g.W, sred, compute_warp);
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
  *smem_bytes = data + 2 * LNP_MAX_STAGES * 8 + 16;
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

template <typename T, int NCH>
static int fwd_launch(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                      long long M, long long F, int relu, DropCfg da, DropCfg dy, cudaStream_t st) {
  LnPipeGeom g;
  size_t smem;
  MMER_TRY(lnp_geometry(M, F, sizeof(T), x ? 2 : 1, &g, &smem));
  static size_t configured = 0;
  auto kern = add_ln_fwd_pipe_kernel<T, NCH>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  kern<<<lnp_grid(g), (g.W + 1) * 32, smem, st>>>((const T*)a, (const T*)x, gamma, beta, (T*)y, stats, g, relu, da, dy);
  MMER_LAUNCH_CHECK("add_ln_fwd_pipe_kernel");
  return 0;
}
template <typename T, int NCH>
static int bwd_launch(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                      const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias, long long M,
                      long long F, int relu, DropCfg da, DropCfg ddy, cudaStream_t st) {
  LnPipeGeom g;
  size_t smem;
  MMER_TRY(lnp_geometry(M, F, sizeof(T), x ? 3 : 2, &g, &smem));
  static size_t configured = 0;
  auto kern = add_ln_bwd_pipe_kernel<T, NCH>;
  MMER_TRY(lnp_set_smem(kern, smem, &configured));
  kern<<<lnp_grid(g), (g.W + 1) * 32, smem, st>>>((const T*)dy, (const T*)a, (const T*)x, stats, gamma, beta, (T*)dz,
                                                   (T*)dap, dgamma, dbeta, dbias, g, relu, da, ddy);
  MMER_LAUNCH_CHECK("add_ln_bwd_pipe_kernel");
  return 0;
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
                    long long F, int dtype, int relu, DropCfg da, DropCfg ddy, cudaStream_t st) {
  if (dtype == MMER_BF16)
    LNP_DISPATCH(F, (bwd_launch<bf16, NCH>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, M, F, relu, da, ddy, st)));
  LNP_DISPATCH(F, (bwd_launch<float, NCH>(dy, x, a, stats, gamma, beta, dz, dap, dgamma, dbeta, dbias, M, F, relu, da, ddy, st)));
}

}  // namespace mmer
