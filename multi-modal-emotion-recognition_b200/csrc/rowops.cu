// Memory-bound row kernels: token assembly (LayerNorm + pos_embed + concat), residual +
// dropout + LayerNorm (+ReLU) forward/backward, masked mean pooling + out_norm, column sums.
// One warp owns one row; a lane owns 8-element (16 B bf16 / 32 B fp32) chunks at
// columns lane*8 + i*256, so every load and store is a fully coalesced 128-bit access.
// Statistics and all reductions are fp32.  Parameter-gradient partials are kept in
// registers across the rows a warp visits, reduced through shared memory per CTA and
// flushed with one fp32 atomic per column per CTA.
#include "common.cuh"

namespace mmer {

// ln_pipe.cu
int add_ln_fwd_pipe(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                    long long M, long long F, int dtype, int relu, DropCfg da, DropCfg dy, cudaStream_t st);
int add_ln_bwd_pipe(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                    const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias, long long M,
                    long long F, int dtype, int relu, DropCfg da, DropCfg ddy, cudaStream_t st, int zin);

int embed_fwd_pipe(const void* pv, const void* pa, const float* gv, const float* bv, const float* ga, const float* ba,
                   const float* pos, void* x0, float* stats, long long B, long long T, long long F, int dtype, DropCfg dc,
                   cudaStream_t st);
int embed_bwd_pipe(const void* dx0, const void* pv, const void* pa, const float* stats, const float* gv, const float* ga,
                   void* dpv, void* dpa, float* dgv, float* dbv, float* dga, float* dba, float* dpos, float* dbias_v,
                   float* dbias_a, long long B, long long T, long long F, int dtype, DropCfg dc, cudaStream_t st);

static constexpr float LN_EPS = 1e-5f;
static constexpr int ROW_WARPS = 8;  // warps per CTA for the row kernels

__device__ __forceinline__ float sum8(const float (&v)[8]) {
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

// reduce per-warp column partials [NCH][8] across the CTA and add them to a global vector
template <int NCH>
__device__ __forceinline__ void flush_cols(float (&part)[NCH][8], float* __restrict__ gout, int F, float* sred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
    if (c < F) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sred[warp * F + c + j] = part[i][j];
    }
  }
  __syncthreads();
  if (gout != nullptr) {
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += sred[w * F + c];
      atomicAdd(gout + c, s);
    }
  }
}

// ---------------------------------------------------------------------------------------
// add + LayerNorm forward
// ---------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ a, const float* __restrict__ gamma,
                  const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ stats, long long M, int F,
                  int relu, DropCfg da, DropCfg dy) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * ROW_WARPS;
  for (long long row = warp_global; row < M; row += nwarps) {
    float z[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        load8(a + off, z[i]);
        if (da.thr) {
          float f[8];
          drop8(da, (uint64_t)off, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] *= f[j];
        }
        if (x != nullptr) {
          float xv[8];
          load8(x + off, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] += xv[j];
        }
        s += sum8(z[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
      }
    }
    const float mean = warp_sum(s) / (float)F;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)F + LN_EPS);
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        float g[8], b[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = (z[i][j] - mean) * rstd * g[j] + b[j];
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

// ---------------------------------------------------------------------------------------
// add + LayerNorm backward
// ---------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
add_ln_bwd_kernel(const T* __restrict__ dyp, const T* __restrict__ x, const T* __restrict__ a,
                  const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                  T* __restrict__ dz, T* __restrict__ dap, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  float* __restrict__ dbias, long long M, int F, int relu, DropCfg da, DropCfg dy) {
  extern __shared__ float sred[];
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * ROW_WARPS;
  float pg[NCH][8], pb[NCH][8], pbias[NCH][8];
  float g[NCH][8], be[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; pbias[i][j] = 0.f; g[i][j] = 0.f; be[i][j] = 0.f; }
    if (c < F) { load8(gamma + c, g[i]); if (relu) load8(beta + c, be[i]); }
  }
  for (long long row = warp_global; row < M; row += nwarps) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float xh[NCH][8], gd[NCH][8], fa[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        float z[8], d[8];
        load8(a + off, z);
        if (da.thr) {
          drop8(da, (uint64_t)off, fa[i]);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] *= fa[i][j];
        }
        if (x != nullptr) {
          float xv[8];
          load8(x + off, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] += xv[j];
        }
        load8(dyp + off, d);
        if (dy.thr) {
          float f[8];
          drop8(dy, (uint64_t)off, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (z[j] - mean) * rstd;
          if (relu && !(xh[i][j] * g[i][j] + be[i][j] > 0.f)) d[j] = 0.f;
          pg[i][j] += d[j] * xh[i][j];
          pb[i][j] += d[j];
          gd[i][j] = d[j] * g[i][j];
          s1 += gd[i][j];
          s2 += gd[i][j] * xh[i][j];
        }
      }
    }
    const float c1 = warp_sum(s1) / (float)F;
    const float c2 = warp_sum(s2) / (float)F;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (gd[i][j] - c1 - xh[i][j] * c2);
        store8(dz + off, o);
        if (da.thr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= fa[i][j];
          if (dap != nullptr) store8(dap + off, o);
        }
        // bias gradient of the Linear that produced `a`: column sum of what is stored
#pragma unroll
        for (int j = 0; j < 8; ++j) pbias[i][j] += round_as<T>(o[j]);
      }
    }
  }
  flush_cols<NCH>(pg, dgamma, F, sred);
  flush_cols<NCH>(pb, dbeta, F, sred);
  if (dbias != nullptr) flush_cols<NCH>(pbias, dbias, F, sred);
}

// ---------------------------------------------------------------------------------------
// token assembly forward / backward.  grid = (blocks, S): every CTA works on one sequence
// position s, so pos_embed[s] and the LayerNorm parameter set (video / audio) are uniform.
// ---------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
embed_fwd_kernel(const T* __restrict__ pv, const T* __restrict__ pa, const float* __restrict__ gv,
                 const float* __restrict__ bv, const float* __restrict__ ga, const float* __restrict__ ba,
                 const float* __restrict__ pos, T* __restrict__ x0, float* __restrict__ stats, int B, int Tn, int F,
                 DropCfg dc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = Tn + 1, s = blockIdx.y;
  const bool is_audio = s == Tn;
  const float* gamma = is_audio ? ga : gv;
  const float* beta = is_audio ? ba : bv;
  float g[NCH][8], be[NCH][8], pe[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
    if (c < F) {
      load8(pos + (long long)s * F + c, pe[i]);
      if (gamma != nullptr) { load8(gamma + c, g[i]); load8(beta + c, be[i]); }
    }
  }
  // EMB_U rows per iteration: all of their 16-byte loads are issued before the first one is consumed, which is
  // what keeps enough bytes in flight for an HBM-bound row kernel
  constexpr int EMB_U = 4;
  const int bstride = gridDim.x * ROW_WARPS;
  for (int b0 = blockIdx.x * ROW_WARPS + warp; b0 < B; b0 += bstride * EMB_U) {
    Raw8<T> zr[EMB_U][NCH];
#pragma unroll
    for (int u = 0; u < EMB_U; ++u) {
      const int b = b0 + u * bstride;
      if (b < B) {
        const T* src = is_audio ? pa + (long long)b * F : pv + ((long long)b * Tn + s) * F;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane * 8 + i * 256;
          if (c < F) zr[u][i].load(src + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < EMB_U; ++u) {
      const int b = b0 + u * bstride;
      if (b >= B) break;
      const long long row = (long long)b * S + s;
      float z[NCH][8];
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane * 8 + i * 256 < F) zr[u][i].get(z[i]);
      float sm = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) { sm += sum8(z[i]); }
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (gamma != nullptr) {
        mean = warp_sum(sm) / (float)F;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane * 8 + i * 256;
          if (c < F) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = z[i][j] - mean; q += d * d; }
          }
        }
        rstd = rsqrtf(warp_sum(q) / (float)F + LN_EPS);
        if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) {
          const long long off = row * F + c;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = (gamma != nullptr ? (z[i][j] - mean) * rstd * g[i][j] + be[i][j] : z[i][j]) + pe[i][j];
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
}

template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
embed_bwd_kernel(const T* __restrict__ dx0, const T* __restrict__ pv, const T* __restrict__ pa,
                 const float* __restrict__ stats, const float* __restrict__ gv, const float* __restrict__ ga,
                 T* __restrict__ dpv, T* __restrict__ dpa, float* __restrict__ dgv, float* __restrict__ dbv,
                 float* __restrict__ dga, float* __restrict__ dba, float* __restrict__ dpos, int B, int Tn, int F,
                 DropCfg dc) {
  extern __shared__ float sred[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = Tn + 1, s = blockIdx.y;
  const bool is_audio = s == Tn;
  const float* gamma = is_audio ? ga : gv;
  float g[NCH][8], pg[NCH][8], pb[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; g[i][j] = 1.f; }
    if (c < F && gamma != nullptr) load8(gamma + c, g[i]);
  }
  // the next row's loads are in flight (as raw 16-byte registers) while the current row is processed
  Raw8<T> nd[NCH], nz[NCH];
  auto prefetch = [&](int b) {
    if (b < B) {
      const long long srow = is_audio ? (long long)b : (long long)b * Tn + s;
      const T* src = (is_audio ? pa : pv) + srow * F;
      const T* dsrc = dx0 + ((long long)b * S + s) * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) { nd[i].load(dsrc + c); nz[i].load(src + c); }
      }
    }
  };
  const int bstride = gridDim.x * ROW_WARPS;
  prefetch(blockIdx.x * ROW_WARPS + warp);
  for (int b = blockIdx.x * ROW_WARPS + warp; b < B; b += bstride) {
    const long long srow = is_audio ? (long long)b : (long long)b * Tn + s;
    T* dst = (is_audio ? dpa : dpv) + srow * F;
    const long long row = (long long)b * S + s;
    float mean = 0.f, rstd = 1.f;
    if (gamma != nullptr) { mean = stats[row * 2]; rstd = stats[row * 2 + 1]; }
    float xh[NCH][8], gd[NCH][8];
    float s1 = 0.f, s2 = 0.f;
    Raw8<T> cd[NCH], cz[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { cd[i] = nd[i]; cz[i] = nz[i]; }
    prefetch(b + bstride);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        const long long off = row * F + c;
        float d[8], z[8];
        cd[i].get(d);
        if (dc.thr) {
          float f[8];
          drop8(dc, (uint64_t)off, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= f[j];
        }
        cz[i].get(z);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pb[i][j] += d[j];  // = dpos contribution = dbeta contribution
          xh[i][j] = (z[j] - mean) * rstd;
          pg[i][j] += d[j] * xh[i][j];
          gd[i][j] = d[j] * g[i][j];
          s1 += gd[i][j];
          s2 += gd[i][j] * xh[i][j];
        }
      }
    }
    if (gamma != nullptr) {
      const float c1 = warp_sum(s1) / (float)F, c2 = warp_sum(s2) / (float)F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (gd[i][j] - c1 - xh[i][j] * c2);
          store8(dst + c, o);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) store8(dst + c, gd[i]);
      }
    }
  }
  if (gamma != nullptr) flush_cols<NCH>(pg, is_audio ? dga : dgv, F, sred);
  // dbeta and dpos[s] share the same column sums
  __syncthreads();
  {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sred[warp * F + c + j] = pb[i][j];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
      float sacc = 0.f;
      for (int w = 0; w < nw; ++w) sacc += sred[w * F + c];
      if (gamma != nullptr) atomicAdd((is_audio ? dba : dbv) + c, sacc);
      if (dpos != nullptr) atomicAdd(dpos + (long long)s * F + c, sacc);
    }
  }
}

// ---------------------------------------------------------------------------------------
// masked mean pooling + out_norm
// ---------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
pool_ln_fwd_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ pooled, T* __restrict__ fused,
                   float* __restrict__ stats, int B, int Tn, int F) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = Tn + 1;
  for (int b = blockIdx.x * ROW_WARPS + warp; b < B; b += gridDim.x * ROW_WARPS) {
    float acc[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    int cnt = 0;
    for (int s = 0; s < S; ++s) {
      const bool valid = (s == Tn) || mask == nullptr || mask[(long long)b * Tn + s] == 0;
      if (!valid) continue;
      ++cnt;
      const T* src = x + ((long long)b * S + s) * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) {
          float v[8];
          load8(src + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
        }
      }
    }
    // reference: sum / clamp(count, 1e-6) with mask, plain mean over S without
    const float inv = 1.f / fmaxf((float)cnt, 1e-6f);
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j] *= inv; sm += acc[i][j]; }
        store8(pooled + (long long)b * F + c, acc[i]);
      }
    }
    if (gamma == nullptr) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) store8(fused + (long long)b * F + c, acc[i]);
      }
      continue;
    }
    const float mean = warp_sum(sm) / (float)F;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = acc[i][j] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)F + LN_EPS);
    if (lane == 0) { stats[b * 2] = mean; stats[b * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        float g[8], be[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (acc[i][j] - mean) * rstd * g[j] + be[j];
        store8(fused + (long long)b * F + c, o);
      }
    }
  }
}

// Long sequences with few samples (cfg4: B = 512, S = 257): a warp per sample leaves most of the machine idle
// (139 us for 135 MB, 0.15 of the HBM peak), so ONE CTA takes a sample: warp w sums rows w, w + 8, ... (four rows of
// loads in flight), the eight partial sums meet in shared memory in warp order (deterministic), warp 0 normalises.
template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
pool_ln_fwd_cta_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float* __restrict__ pooled, T* __restrict__ fused,
                       float* __restrict__ stats, int B, int Tn, int F) {
  extern __shared__ float pool_sred[];   // [ROW_WARPS][F] partial sums, then [ROW_WARPS] valid-row counts
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = Tn + 1;
  int* scnt = reinterpret_cast<int*>(pool_sred + ROW_WARPS * F);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float acc[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    int cnt = 0;
    constexpr int U = 4;
    for (int s0 = warp; s0 < S; s0 += ROW_WARPS * U) {
      Raw8<T> r[U][NCH];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u * ROW_WARPS;
        ok[u] = s < S && ((s == Tn) || mask == nullptr || mask[(long long)b * Tn + s] == 0);
        if (ok[u]) {
          const T* src = x + ((long long)b * S + s) * F;
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            if (lane * 8 + i * 256 < F) r[u][i].load(src + lane * 8 + i * 256);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        ++cnt;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          if (lane * 8 + i * 256 < F) {
            float v[8];
            r[u][i].get(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
          }
      }
    }
    __syncthreads();   // the previous sample's partial sums have been consumed
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        *reinterpret_cast<float4*>(pool_sred + warp * F + c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(pool_sred + warp * F + c + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
    if (lane == 0) scnt[warp] = cnt;
    __syncthreads();
    if (warp != 0) continue;
    cnt = 0;
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) cnt += scnt[w];
    const float inv = 1.f / fmaxf((float)cnt, 1e-6f);
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int w = 0; w < ROW_WARPS; ++w) {
          float v[8];
          load8(pool_sred + w * F + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j] *= inv; sm += acc[i][j]; }
        store8(pooled + (long long)b * F + c, acc[i]);
      }
    }
    if (gamma == nullptr) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) store8(fused + (long long)b * F + c, acc[i]);
      }
      continue;
    }
    const float mean = warp_sum(sm) / (float)F;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = acc[i][j] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)F + LN_EPS);
    if (lane == 0) { stats[b * 2] = mean; stats[b * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        float g[8], be[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (acc[i][j] - mean) * rstd * g[j] + be[j];
        store8(fused + (long long)b * F + c, o);
      }
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
pool_ln_bwd_kernel(const T* __restrict__ dfused, const float* __restrict__ pooled, const float* __restrict__ stats,
                   const float* __restrict__ gamma, const uint8_t* __restrict__ mask, T* __restrict__ dx,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int Tn, int F) {
  extern __shared__ float sred[];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = Tn + 1;
  float pg[NCH][8], pb[NCH][8], g[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane * 8 + i * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; g[i][j] = 1.f; }
    if (c < F && gamma != nullptr) load8(gamma + c, g[i]);
  }
  for (int b = blockIdx.x * ROW_WARPS + warp; b < B; b += gridDim.x * ROW_WARPS) {
    float dp[NCH][8];
    float mean = 0.f, rstd = 1.f;
    if (gamma != nullptr) { mean = stats[b * 2]; rstd = stats[b * 2 + 1]; }
    float xh[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane * 8 + i * 256;
      if (c < F) {
        float d[8], z[8];
        load8(dfused + (long long)b * F + c, d);
        load8(pooled + (long long)b * F + c, z);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (z[j] - mean) * rstd;
          pg[i][j] += d[j] * xh[i][j];
          pb[i][j] += d[j];
          dp[i][j] = d[j] * g[i][j];
          s1 += dp[i][j];
          s2 += dp[i][j] * xh[i][j];
        }
      }
    }
    if (gamma != nullptr) {
      const float c1 = warp_sum(s1) / (float)F, c2 = warp_sum(s2) / (float)F;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dp[i][j] = rstd * (dp[i][j] - c1 - xh[i][j] * c2);
    }
    int cnt = 0;
    for (int s = 0; s < S; ++s)
      cnt += ((s == Tn) || mask == nullptr || mask[(long long)b * Tn + s] == 0) ? 1 : 0;
    const float inv = 1.f / fmaxf((float)cnt, 1e-6f);
    for (int s = 0; s < S; ++s) {
      const bool valid = (s == Tn) || mask == nullptr || mask[(long long)b * Tn + s] == 0;
      T* dst = dx + ((long long)b * S + s) * F;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = lane * 8 + i * 256;
        if (c < F) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = valid ? dp[i][j] * inv : 0.f;
          store8(dst + c, o);
        }
      }
    }
  }
  if (gamma != nullptr) {
    flush_cols<NCH>(pg, dgamma, F, sred);
    flush_cols<NCH>(pb, dbeta, F, sred);
  }
}

// ---------------------------------------------------------------------------------------
// column sums: out[n] += sum_m x[m,n].  block (32 lanes x 8 warps) covers 256 columns.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long M, int N, long long ldx) {
  __shared__ float sred[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    for (long long r = (long long)blockIdx.y * 8 + warp; r < M; r += (long long)gridDim.y * 8) {
      float v[8];
      load8(x + r * ldx + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sred[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sred[w][threadIdx.x];
    atomicAdd(out + cc, s);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    load8(src + i, v);
    store8(dst + i, v);
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
  }
}
__global__ void cast_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    load8(src + i, v);
    store8(dst + i, v);
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __bfloat162float(src[k]);
  }
}

// ---------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------
static int nch_for(long long F) { return F <= 256 ? 1 : F <= 512 ? 2 : F <= 1024 ? 4 : 8; }
static int row_grid(long long rows) {
  long long want = (rows + ROW_WARPS - 1) / ROW_WARPS;
  long long cap = (long long)sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
static int bwd_grid(long long rows) {
  long long want = (rows + ROW_WARPS - 1) / ROW_WARPS;
  long long cap = (long long)sm_count() * 2;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

#define DISPATCH_NCH(F, ...)                                   \
  switch (nch_for(F)) {                                        \
    case 1: { constexpr int NCH = 1; __VA_ARGS__; } break;     \
    case 2: { constexpr int NCH = 2; __VA_ARGS__; } break;     \
    case 4: { constexpr int NCH = 4; __VA_ARGS__; } break;     \
    default: { constexpr int NCH = 8; __VA_ARGS__; } break;    \
  }

#define CHECK_ROW_SHAPE(F)                                                                               \
  MMER_CHECK_ARG((F) > 0 && (F) % 8 == 0 && (F) <= 2048, "row width %lld must be a multiple of 8, <= 2048", \
                 (long long)(F))

template <typename T>
static int add_ln_fwd_t(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                        long long M, long long F, int relu, DropCfg da, DropCfg dy, cudaStream_t st) {
  DISPATCH_NCH(F, (add_ln_fwd_kernel<T, NCH><<<row_grid(M), ROW_WARPS * 32, 0, st>>>(
                      (const T*)x, (const T*)a, gamma, beta, (T*)y, stats, M, (int)F, relu, da, dy)));
  MMER_LAUNCH_CHECK("add_ln_fwd_kernel");
  return 0;
}
template <typename T>
static int add_ln_bwd_t(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                        const float* beta, void* dz, void* dap, float* dgamma, float* dbeta, float* dbias,
                        long long M, long long F, int relu, DropCfg da, DropCfg ddy, cudaStream_t st) {
  const size_t sm = (size_t)ROW_WARPS * F * sizeof(float);
  DISPATCH_NCH(F, (add_ln_bwd_kernel<T, NCH><<<bwd_grid(M), ROW_WARPS * 32, sm, st>>>(
                      (const T*)dy, (const T*)x, (const T*)a, stats, gamma, beta, (T*)dz, (T*)dap, dgamma, dbeta,
                      dbias, M, (int)F, relu, da, ddy)));
  MMER_LAUNCH_CHECK("add_ln_bwd_kernel");
  return 0;
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_add_ln_fwd(const void* x, const void* a, const float* gamma, const float* beta, void* y, float* stats,
                    int64_t M, int64_t F, int dtype, int relu, float drop_a_p, uint32_t site_a, float drop_y_p,
                    uint32_t site_y, uint64_t seed, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(a && gamma && beta && y && stats, "add_ln_fwd: null pointer");
  if (M <= 0) return 0;
  DropCfg da = make_drop(drop_a_p, seed, site_a), dy = make_drop(drop_y_p, seed, site_y);
  cudaStream_t st = (cudaStream_t)stream;
  return add_ln_fwd_pipe(x, a, gamma, beta, y, stats, M, F, dtype, relu, da, dy, st);
}

int mmer_add_ln_bwd(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                    const float* beta, void* dz, void* da, float* dgamma, float* dbeta, float* dbias, int64_t M,
                    int64_t F, int dtype, int relu, float drop_a_p, uint32_t site_a, float drop_y_p,
                    uint32_t site_y, uint64_t seed, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(dy && a && gamma && dz && stats, "add_ln_bwd: null pointer");
  MMER_CHECK_ARG(!relu || beta != nullptr, "add_ln_bwd: relu needs beta");
  if (M <= 0) return 0;
  DropCfg dca = make_drop(drop_a_p, seed, site_a), dcy = make_drop(drop_y_p, seed, site_y);
  cudaStream_t st = (cudaStream_t)stream;
  return add_ln_bwd_pipe(dy, x, a, stats, gamma, beta, dz, da, dgamma, dbeta, dbias, M, F, dtype, relu, dca, dcy, st, 0);
}

int mmer_add_ln_bwd_z(const void* dy, const void* z, const float* stats, const float* gamma, void* dz, void* da,
                      float* dgamma, float* dbeta, float* dbias, int64_t M, int64_t F, int dtype, float drop_a_p,
                      uint32_t site_a, uint64_t seed, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(dy && z && gamma && dz && stats, "add_ln_bwd_z: null pointer");
  if (M <= 0) return 0;
  DropCfg dca = make_drop(drop_a_p, seed, site_a), none = make_drop(0.f, 0, 0);
  return add_ln_bwd_pipe(dy, nullptr, z, stats, gamma, nullptr, dz, da, dgamma, dbeta, dbias, M, F, dtype, 0, dca, none,
                         (cudaStream_t)stream, 1);
}

int mmer_embed_fwd(const void* pv, const void* pa, const float* gv, const float* bv, const float* ga, const float* ba,
                   const float* pos, void* x0, float* stats, int64_t B, int64_t T, int64_t F, int dtype, float drop_p,
                   uint64_t seed, uint32_t site, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(pv && pa && pos && x0, "embed_fwd: null pointer");
  MMER_CHECK_ARG((gv == nullptr) == (ga == nullptr), "embed_fwd: both or neither LayerNorm");
  if (B <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (gv != nullptr && stats != nullptr)   // LayerNorm variant: bulk-copy pipelined kernel (ln_pipe.cu)
    return embed_fwd_pipe(pv, pa, gv, bv, ga, ba, pos, x0, stats, B, T, F, dtype, dc, st);
  int gx = (int)((B + ROW_WARPS - 1) / ROW_WARPS);
  int cap = (sm_count() * 8) / (int)(T + 1) + 1;
  if (gx > cap) gx = cap;
  dim3 grid(gx, (unsigned)(T + 1));
  if (dtype == MMER_BF16) {
    DISPATCH_NCH(F, (embed_fwd_kernel<bf16, NCH><<<grid, ROW_WARPS * 32, 0, st>>>(
                        (const bf16*)pv, (const bf16*)pa, gv, bv, ga, ba, pos, (bf16*)x0, stats, (int)B, (int)T, (int)F, dc)));
  } else {
    DISPATCH_NCH(F, (embed_fwd_kernel<float, NCH><<<grid, ROW_WARPS * 32, 0, st>>>(
                        (const float*)pv, (const float*)pa, gv, bv, ga, ba, pos, (float*)x0, stats, (int)B, (int)T, (int)F, dc)));
  }
  MMER_LAUNCH_CHECK("embed_fwd_kernel");
  return 0;
}

int mmer_embed_bwd(const void* dx0, const void* pv, const void* pa, const float* stats, const float* gv,
                   const float* ga, void* dpv, void* dpa, float* dgv, float* dbv, float* dga, float* dba, float* dpos,
                   float* dbias_v, float* dbias_a, int64_t B, int64_t T, int64_t F, int dtype, float drop_p, uint64_t seed,
                   uint32_t site, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(dx0 && pv && pa && dpv && dpa, "embed_bwd: null pointer");
  if (B <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (gv != nullptr && ga != nullptr && stats != nullptr && dgv && dga)
    return embed_bwd_pipe(dx0, pv, pa, stats, gv, ga, dpv, dpa, dgv, dbv, dga, dba, dpos, dbias_v, dbias_a, B, T, F, dtype,
                          dc, st);
  MMER_CHECK_ARG(dbias_v == nullptr && dbias_a == nullptr, "embed_bwd: bias-gradient outputs need the LayerNorm variant");
  int gx = (int)((B + ROW_WARPS - 1) / ROW_WARPS);
  int cap = (sm_count() * 2) / (int)(T + 1) + 1;
  if (gx > cap) gx = cap;
  dim3 grid(gx, (unsigned)(T + 1));
  const size_t sm = (size_t)ROW_WARPS * F * sizeof(float);
  if (dtype == MMER_BF16) {
    DISPATCH_NCH(F, (embed_bwd_kernel<bf16, NCH><<<grid, ROW_WARPS * 32, sm, st>>>(
                        (const bf16*)dx0, (const bf16*)pv, (const bf16*)pa, stats, gv, ga, (bf16*)dpv, (bf16*)dpa, dgv,
                        dbv, dga, dba, dpos, (int)B, (int)T, (int)F, dc)));
  } else {
    DISPATCH_NCH(F, (embed_bwd_kernel<float, NCH><<<grid, ROW_WARPS * 32, sm, st>>>(
                        (const float*)dx0, (const float*)pv, (const float*)pa, stats, gv, ga, (float*)dpv, (float*)dpa,
                        dgv, dbv, dga, dba, dpos, (int)B, (int)T, (int)F, dc)));
  }
  MMER_LAUNCH_CHECK("embed_bwd_kernel");
  return 0;
}

int mmer_pool_ln_fwd(const void* x, const uint8_t* mask, const float* gamma, const float* beta, float* pooled,
                     void* fused, float* stats, int64_t B, int64_t T, int64_t F, int dtype, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(x && pooled && fused, "pool_ln_fwd: null pointer");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t le = cudaSuccess;
  // few samples with long sequences: one CTA per sample instead of one warp (see pool_ln_fwd_cta_kernel)
  if (T + 1 >= 64 && B < (int64_t)sm_count() * ROW_WARPS * 2 && F <= 1024) {
    const size_t sm_bytes = (size_t)ROW_WARPS * F * sizeof(float) + ROW_WARPS * sizeof(int);
    const long long cap = (long long)sm_count() * 4;
    const dim3 grid((unsigned)(B < cap ? B : cap));
    if (dtype == MMER_BF16) {
      DISPATCH_NCH(F, (le = launch_dep(pool_ln_fwd_cta_kernel<bf16, NCH>, grid, dim3(ROW_WARPS * 32), sm_bytes, st, 1,
                                       (const bf16*)x, mask, gamma, beta, pooled, (bf16*)fused, stats, (int)B, (int)T, (int)F)));
    } else {
      DISPATCH_NCH(F, (le = launch_dep(pool_ln_fwd_cta_kernel<float, NCH>, grid, dim3(ROW_WARPS * 32), sm_bytes, st, 1,
                                       (const float*)x, mask, gamma, beta, pooled, (float*)fused, stats, (int)B, (int)T, (int)F)));
    }
    if (le != cudaSuccess) return cuda_fail(le, "launch(pool_ln_fwd_cta)");
    MMER_LAUNCH_CHECK("pool_ln_fwd_cta_kernel");
    return 0;
  }
  if (dtype == MMER_BF16) {
    DISPATCH_NCH(F, (le = launch_dep(pool_ln_fwd_kernel<bf16, NCH>, dim3(row_grid(B)), dim3(ROW_WARPS * 32), 0, st, 1,
                                     (const bf16*)x, mask, gamma, beta, pooled, (bf16*)fused, stats, (int)B, (int)T, (int)F)));
  } else {
    DISPATCH_NCH(F, (le = launch_dep(pool_ln_fwd_kernel<float, NCH>, dim3(row_grid(B)), dim3(ROW_WARPS * 32), 0, st, 1,
                                     (const float*)x, mask, gamma, beta, pooled, (float*)fused, stats, (int)B, (int)T, (int)F)));
  }
  if (le != cudaSuccess) return cuda_fail(le, "launch(pool_ln_fwd)");
  MMER_LAUNCH_CHECK("pool_ln_fwd_kernel");
  return 0;
}

int mmer_pool_ln_bwd(const void* dfused, const float* pooled, const float* stats, const float* gamma,
                     const uint8_t* mask, void* dx, float* dgamma, float* dbeta, int64_t B, int64_t T, int64_t F,
                     int dtype, void* stream) {
  CHECK_ROW_SHAPE(F);
  MMER_CHECK_ARG(dfused && pooled && dx, "pool_ln_bwd: null pointer");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t le = cudaSuccess;
  const size_t sm = (size_t)ROW_WARPS * F * sizeof(float);
  if (dtype == MMER_BF16) {
    DISPATCH_NCH(F, (le = launch_dep(pool_ln_bwd_kernel<bf16, NCH>, dim3(bwd_grid(B)), dim3(ROW_WARPS * 32), sm, st, 1,
                                     (const bf16*)dfused, pooled, stats, gamma, mask, (bf16*)dx, dgamma, dbeta, (int)B, (int)T,
                                     (int)F)));
  } else {
    DISPATCH_NCH(F, (le = launch_dep(pool_ln_bwd_kernel<float, NCH>, dim3(bwd_grid(B)), dim3(ROW_WARPS * 32), sm, st, 1,
                                     (const float*)dfused, pooled, stats, gamma, mask, (float*)dx, dgamma, dbeta, (int)B, (int)T,
                                     (int)F)));
  }
  if (le != cudaSuccess) return cuda_fail(le, "launch(pool_ln_bwd)");
  MMER_LAUNCH_CHECK("pool_ln_bwd_kernel");
  return 0;
}

int mmer_colsum(const void* x, float* out, int64_t M, int64_t N, int64_t ldx, int dtype, void* stream) {
  MMER_CHECK_ARG(x && out && N > 0 && N % 8 == 0 && ldx % 8 == 0, "colsum: N and ldx must be multiples of 8");
  if (M <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int gx = (int)((N + 255) / 256);
  int gy = (sm_count() * 4) / gx;
  long long maxy = (M + 7) / 8;
  if (gy > maxy) gy = (int)maxy;
  if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  if (dtype == MMER_BF16) colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, out, M, (int)N, ldx);
  else colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, M, (int)N, ldx);
  MMER_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

int mmer_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  MMER_CHECK_ARG(src && dst, "cast_bf16: null pointer");
  long long nt = (n + 7) / 8;
  cast_bf16_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  MMER_LAUNCH_CHECK("cast_bf16_kernel");
  return 0;
}
int mmer_cast_f32(const void* src, float* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  MMER_CHECK_ARG(src && dst, "cast_f32: null pointer");
  long long nt = (n + 7) / 8;
  cast_f32_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  MMER_LAUNCH_CHECK("cast_f32_kernel");
  return 0;
}

}  // extern "C"
