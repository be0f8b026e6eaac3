// Long-sequence attention (S = T+1 > 32, e.g. the T=256 interpretability configuration, S = 257).
//
// One CTA owns one (sample, head).  Two fp32 matrices of the head ([S][D], rows padded to D+4 so
// that "one row per lane" float4 reads are bank-conflict free) live in shared memory; each warp
// takes one row of the "outer" side at a time:
//   forward           K, V resident; a warp takes query i: scores against every key (keys strided over
//                     lanes), warp-wide softmax, then O_i = sum_j p_j V_j with two columns per lane
//   backward, pass A  K, V resident; per query i: p_i, dP_i = dO_i V^T, dS_i, dQ_i = dS_i K; the row
//                     statistics (max, 1/sum, sum_j dP*P) are kept in shared memory
//   backward, pass B  Q, dO resident (same buffers); a warp takes key j: recomputes column j of P and
//                     dS from the saved statistics, dK_j = sum_i dS_ij Q_i, dV_j = sum_i Pd_ij dO_i
// Nothing but the packed QKV tensor is kept from the forward pass; dropout masks are regenerated
// from (seed, site, element index) exactly as in the short-sequence kernels.
#include "common.cuh"

namespace mmer {

static constexpr int GA_WARPS = 8;
static constexpr int GA_MAX_T = 13;  // keys per lane: S <= 32 * 13 = 416

template <int D> struct GaSmem {
  static constexpr int RP = D + 4;
  // floats: two [S][RP] matrices, per-warp vector [D] and two per-warp [S4] buffers, 3 stats rows
  __host__ __device__ static size_t floats(int S) {
    const int S4 = (S + 3) & ~3;
    return (size_t)2 * S * RP + (size_t)GA_WARPS * (D + 2 * S4) + (size_t)3 * S4;
  }
};

template <typename T, int D>
__device__ __forceinline__ void ga_load_matrix(const T* __restrict__ g, long long ld, float* s, int S) {
  constexpr int RP = D + 4, CH = D / 8;
  for (int e = threadIdx.x; e < S * CH; e += blockDim.x) {
    const int r = e / CH, c = (e % CH) * 8;
    float v[8];
    load8(g + (long long)r * ld + c, v);
    *reinterpret_cast<float4*>(s + r * RP + c) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(s + r * RP + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// vec[D] (per-warp smem) <- one global row
template <typename T, int D>
__device__ __forceinline__ void ga_load_vec(const T* __restrict__ g, float* vec, int lane) {
  for (int c = lane; c < D; c += 32) vec[c] = to_f(g[c]);
}

// dot of the broadcast vector with row j of a padded matrix
template <int D>
__device__ __forceinline__ float ga_dot(const float* __restrict__ vec, const float* __restrict__ row) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int d = 0; d < D / 4; ++d) {
    const float4 x = *reinterpret_cast<const float4*>(vec + d * 4);
    const float4 y = *reinterpret_cast<const float4*>(row + d * 4);
    a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1); a2 = fmaf(x.z, y.z, a2); a3 = fmaf(x.w, y.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// acc[c] = sum_j w[j] * Mx[j][lane*CPL + c]   (w: S4 floats in smem, zero padded)
template <int D>
__device__ __forceinline__ void ga_weighted_cols(const float* __restrict__ w, const float* __restrict__ Mx, int S,
                                                 int lane, float (&acc)[D / 32]) {
  constexpr int RP = D + 4, CPL = D / 32;
#pragma unroll
  for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
  const float* col = Mx + lane * CPL;
  int j = 0;
  for (; j + 4 <= S; j += 4) {
    const float4 ww = *reinterpret_cast<const float4*>(w + j);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      acc[c] = fmaf(ww.x, col[(j + 0) * RP + c], acc[c]);
      acc[c] = fmaf(ww.y, col[(j + 1) * RP + c], acc[c]);
      acc[c] = fmaf(ww.z, col[(j + 2) * RP + c], acc[c]);
      acc[c] = fmaf(ww.w, col[(j + 3) * RP + c], acc[c]);
    }
  }
  for (; j < S; ++j) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = fmaf(w[j], col[j * RP + c], acc[c]);
  }
}

template <typename T, int CPL>
__device__ __forceinline__ void ga_store_cols(T* __restrict__ g, const float (&acc)[CPL]) {
#pragma unroll
  for (int c = 0; c < CPL; ++c) g[c] = from_f<T>(acc[c]);
}

template <typename T, int D>
__global__ void __launch_bounds__(GA_WARPS * 32)
mha_fwd_generic_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, T* __restrict__ out,
                       float* __restrict__ probs, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RP = D + 4, CPL = D / 32;
  const int S = Tn + 1, S4 = (S + 3) & ~3, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = smem;
  float* Vs = Ks + (size_t)S * RP;
  float* vec = Vs + (size_t)S * RP + warp * (D + 2 * S4);
  float* pbuf = vec + D;
  const float scale = rsqrtf((float)D);
  for (long long bh = blockIdx.x; bh < (long long)B * H; bh += gridDim.x) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    __syncthreads();
    ga_load_matrix<T, D>(base + F, 3 * F, Ks, S);
    ga_load_matrix<T, D>(base + 2 * F, 3 * F, Vs, S);
    __syncthreads();
    for (int i = warp; i < S; i += GA_WARPS) {
      ga_load_vec<T, D>(base + (long long)i * 3 * F, vec, lane);
      __syncwarp();
      float s[GA_MAX_T];
      float m = -INFINITY;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int j = lane + 32 * t;
        s[t] = -INFINITY;
        if (j < S) {
          const bool ok = j == Tn || mask == nullptr || mask[(long long)b * Tn + j] == 0;
          if (ok) s[t] = ga_dot<D>(vec, Ks + (size_t)j * RP) * scale;
        }
        m = fmaxf(m, s[t]);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        s[t] = (lane + 32 * t < S) ? __expf(s[t] - m) : 0.f;
        sum += s[t];
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int j = lane + 32 * t;
        if (j < S4) {
          float p = s[t] * inv;
          if (j < S) {
            if (probs) probs[bh * S * S + (long long)i * S + j] = p;
            if (dc.thr) p *= drop1(dc, att_drop_index(bh * S + i, j, att_drop_stride(S)));
          } else {
            p = 0.f;
          }
          pbuf[j] = p;
        }
      }
      __syncwarp();
      float acc[CPL];
      ga_weighted_cols<D>(pbuf, Vs, S, lane, acc);
      ga_store_cols<T, CPL>(out + ((long long)b * S + i) * F + h * D + lane * CPL, acc);
      __syncwarp();
    }
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(GA_WARPS * 32)
mha_bwd_generic_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, const T* __restrict__ dout,
                       T* __restrict__ dqkv, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RP = D + 4, CPL = D / 32;
  const int S = Tn + 1, S4 = (S + 3) & ~3, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* M1 = smem;                       // K, then Q
  float* M2 = M1 + (size_t)S * RP;        // V, then dO
  float* wbase = M2 + (size_t)S * RP;
  float* vec = wbase + warp * (D + 2 * S4);
  float* buf1 = vec + D;
  float* buf2 = buf1 + S4;
  float* st_m = wbase + GA_WARPS * (D + 2 * S4);
  float* st_inv = st_m + S4;
  float* st_dot = st_inv + S4;
  const float scale = rsqrtf((float)D);
  for (long long bh = blockIdx.x; bh < (long long)B * H; bh += gridDim.x) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    const T* dobase = dout + (long long)b * S * F + h * D;
    T* dbase = dqkv + (long long)b * S * 3 * F + h * D;
    __syncthreads();
    ga_load_matrix<T, D>(base + F, 3 * F, M1, S);
    ga_load_matrix<T, D>(base + 2 * F, 3 * F, M2, S);
    __syncthreads();
    // ---------------- pass A: one query per warp
    for (int i = warp; i < S; i += GA_WARPS) {
      float s[GA_MAX_T], dp[GA_MAX_T];
      ga_load_vec<T, D>(base + (long long)i * 3 * F, vec, lane);
      __syncwarp();
      float m = -INFINITY;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int j = lane + 32 * t;
        s[t] = -INFINITY;
        if (j < S) {
          const bool ok = j == Tn || mask == nullptr || mask[(long long)b * Tn + j] == 0;
          if (ok) s[t] = ga_dot<D>(vec, M1 + (size_t)j * RP) * scale;
        }
        m = fmaxf(m, s[t]);
      }
      m = warp_max(m);
      __syncwarp();
      ga_load_vec<T, D>(dobase + (long long)i * F, vec, lane);
      __syncwarp();
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int j = lane + 32 * t;
        s[t] = (j < S) ? __expf(s[t] - m) : 0.f;
        sum += s[t];
        dp[t] = 0.f;
        if (j < S) {
          dp[t] = ga_dot<D>(vec, M2 + (size_t)j * RP);
          if (dc.thr) dp[t] *= drop1(dc, att_drop_index(bh * S + i, j, att_drop_stride(S)));
        }
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      float dot = 0.f;
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) dot = fmaf(s[t] * inv, dp[t], dot);
      dot = warp_sum(dot);
      if (lane == 0) { st_m[i] = m; st_inv[i] = inv; st_dot[i] = dot; }
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int j = lane + 32 * t;
        if (j < S4) buf1[j] = (j < S) ? s[t] * inv * (dp[t] - dot) * scale : 0.f;
      }
      __syncwarp();
      float acc[CPL];
      ga_weighted_cols<D>(buf1, M1, S, lane, acc);   // dQ_i = sum_j dS_ij K_j
      ga_store_cols<T, CPL>(dbase + (long long)i * 3 * F + lane * CPL, acc);
      __syncwarp();
    }
    __syncthreads();
    ga_load_matrix<T, D>(base, 3 * F, M1, S);     // Q
    ga_load_matrix<T, D>(dobase, F, M2, S);       // dO
    __syncthreads();
    // ---------------- pass B: one key per warp
    for (int j = warp; j < S; j += GA_WARPS) {
      const bool ok = j == Tn || mask == nullptr || mask[(long long)b * Tn + j] == 0;
      T* dk = dbase + (long long)j * 3 * F + F + lane * CPL;
      T* dv = dbase + (long long)j * 3 * F + 2 * F + lane * CPL;
      float acc[CPL];
      if (!ok) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
        ga_store_cols<T, CPL>(dk, acc);
        ga_store_cols<T, CPL>(dv, acc);
        continue;
      }
      float p[GA_MAX_T];
      ga_load_vec<T, D>(base + (long long)j * 3 * F + F, vec, lane);   // k_j
      __syncwarp();
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int i = lane + 32 * t;
        p[t] = 0.f;
        if (i < S) p[t] = __expf(ga_dot<D>(vec, M1 + (size_t)i * RP) * scale - st_m[i]) * st_inv[i];
      }
      __syncwarp();
      ga_load_vec<T, D>(base + (long long)j * 3 * F + 2 * F, vec, lane);  // v_j
      __syncwarp();
#pragma unroll
      for (int t = 0; t < GA_MAX_T; ++t) {
        const int i = lane + 32 * t;
        if (i < S4) {
          float ds = 0.f, pd = 0.f;
          if (i < S) {
            float f = 1.f;
            if (dc.thr) f = drop1(dc, att_drop_index(bh * S + i, j, att_drop_stride(S)));
            const float dpv = ga_dot<D>(vec, M2 + (size_t)i * RP) * f;
            ds = p[t] * (dpv - st_dot[i]) * scale;
            pd = p[t] * f;
          }
          buf1[i] = ds;
          buf2[i] = pd;
        }
      }
      __syncwarp();
      ga_weighted_cols<D>(buf1, M1, S, lane, acc);   // dK_j = sum_i dS_ij Q_i
      ga_store_cols<T, CPL>(dk, acc);
      ga_weighted_cols<D>(buf2, M2, S, lane, acc);   // dV_j = sum_i Pd_ij dO_i
      ga_store_cols<T, CPL>(dv, acc);
      __syncwarp();
    }
  }
}

template <typename K>
static int ga_configure(K kern, size_t smem) {
  int dev = 0, maxs = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if ((long long)smem > (long long)maxs) {
    set_error("mha: sequence too long for the shared-memory resident kernel (%lld bytes needed, %d available)",
              (long long)smem, maxs);
    return MMER_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha_generic)");
  return 0;
}

template <typename T, int D>
static int ga_fwd(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, DropCfg dc,
                  cudaStream_t st) {
  const size_t smem = GaSmem<D>::floats(Tn + 1) * sizeof(float);
  auto kern = mha_fwd_generic_kernel<T, D>;
  MMER_TRY(ga_configure(kern, smem));
  const long long want = (long long)B * H, cap = (long long)sm_count() * 4;
  kern<<<(unsigned)(want < cap ? want : cap), GA_WARPS * 32, smem, st>>>((const T*)qkv, mask, (T*)out, probs, B, Tn, H, dc);
  MMER_LAUNCH_CHECK("mha_fwd_generic_kernel");
  return 0;
}
template <typename T, int D>
static int ga_bwd(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H, DropCfg dc,
                  cudaStream_t st) {
  const size_t smem = GaSmem<D>::floats(Tn + 1) * sizeof(float);
  auto kern = mha_bwd_generic_kernel<T, D>;
  MMER_TRY(ga_configure(kern, smem));
  const long long want = (long long)B * H, cap = (long long)sm_count() * 4;
  kern<<<(unsigned)(want < cap ? want : cap), GA_WARPS * 32, smem, st>>>((const T*)qkv, mask, (const T*)dout, (T*)dqkv, B, Tn,
                                                                        H, dc);
  MMER_LAUNCH_CHECK("mha_bwd_generic_kernel");
  return 0;
}

int mha_fwd_generic(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H,
                    int64_t d, int dtype, DropCfg dc, cudaStream_t st) {
  MMER_CHECK_ARG(T + 1 <= 32 * GA_MAX_T, "mha_fwd: sequences longer than %d tokens are not supported (T=%lld)",
                 32 * GA_MAX_T, (long long)T);
  if (dtype == MMER_BF16)
    return d == 64 ? ga_fwd<bf16, 64>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st)
                   : ga_fwd<bf16, 32>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
  return d == 64 ? ga_fwd<float, 64>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st)
                 : ga_fwd<float, 32>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
}
int mha_bwd_generic(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int64_t B, int64_t T,
                    int64_t H, int64_t d, int dtype, DropCfg dc, cudaStream_t st) {
  MMER_CHECK_ARG(T + 1 <= 32 * GA_MAX_T, "mha_bwd: sequences longer than %d tokens are not supported (T=%lld)",
                 32 * GA_MAX_T, (long long)T);
  if (dtype == MMER_BF16)
    return d == 64 ? ga_bwd<bf16, 64>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st)
                   : ga_bwd<bf16, 32>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st);
  return d == 64 ? ga_bwd<float, 64>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st)
                 : ga_bwd<float, 32>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st);
}

}  // namespace mmer
