// Multi-head self-attention over the short token sequence [T video tokens ; audio token].
//
// Short-sequence path (S = T+1 <= 32, the benchmark shape is S = 17): one warp owns one
// (sample, head).  Q/K/V (and dO in backward) are staged in shared memory with coalesced
// 128-bit loads of the packed in_proj output; scores, softmax, dropout and the weighted sum
// stay on chip; results leave through shared memory as coalesced 128-bit stores.
// The backward kernel recomputes the probabilities, so nothing but the packed QKV tensor is
// kept from the forward pass.  Key padding: masked keys get probability exactly 0.
#include "common.cuh"

namespace mmer {

static constexpr int ATT_WARPS = 4;

template <int D> struct AttSmem {
  // per warp, in floats
  static constexpr int KP = D + 4;  // padded K row: conflict-free float4 reads with one key per lane
  __host__ __device__ static int per_warp_fwd(int S) { return (S * D * 2 + S * KP + S * (S | 1) + 3) & ~3; }
  __host__ __device__ static int per_warp_bwd(int S) { return (S * D * 3 + S * KP + 2 * S * (S | 1) + 3) & ~3; }
};

// cooperative load of a [S][D] head slice (row stride ld elements) into smem (row stride rs floats)
template <typename T, int D>
__device__ __forceinline__ void load_head(const T* __restrict__ g, long long ld, float* s, int rs, int S, int lane) {
  constexpr int LPR = D / 8;           // lanes per row
  constexpr int RPP = 32 / LPR;        // rows per pass
  const int c = (lane % LPR) * 8;
  for (int r0 = 0; r0 < S; r0 += RPP) {
    const int r = r0 + lane / LPR;
    if (r < S) {
      float v[8];
      load8(g + (long long)r * ld + c, v);
      *reinterpret_cast<float4*>(s + r * rs + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(s + r * rs + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}
template <typename T, int D>
__device__ __forceinline__ void store_head(T* __restrict__ g, long long ld, const float* s, int rs, int S, int lane) {
  constexpr int LPR = D / 8;
  constexpr int RPP = 32 / LPR;
  const int c = (lane % LPR) * 8;
  for (int r0 = 0; r0 < S; r0 += RPP) {
    const int r = r0 + lane / LPR;
    if (r < S) {
      float v[8];
      const float4 a = *reinterpret_cast<const float4*>(s + r * rs + c);
      const float4 b = *reinterpret_cast<const float4*>(s + r * rs + c + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      store8(g + (long long)r * ld + c, v);
    }
  }
}

// scores[i][j] = scale * q_i . k_j for the key owned by this lane; masked keys -> -inf
template <int D>
__device__ __forceinline__ void scores_phase(const float* Qs, const float* Ks, float* Ps, int S, int SP, int lane,
                                             bool key_ok, float scale) {
  constexpr int KP = AttSmem<D>::KP;
  if (lane < S) {
    float4 kreg[D / 4];
#pragma unroll
    for (int d = 0; d < D / 4; ++d) kreg[d] = *reinterpret_cast<const float4*>(Ks + lane * KP + d * 4);
    for (int i = 0; i < S; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D / 4; ++d) {
        const float4 q = *reinterpret_cast<const float4*>(Qs + i * D + d * 4);
        acc = fmaf(q.x, kreg[d].x, acc); acc = fmaf(q.y, kreg[d].y, acc);
        acc = fmaf(q.z, kreg[d].z, acc); acc = fmaf(q.w, kreg[d].w, acc);
      }
      Ps[i * SP + lane] = key_ok ? acc * scale : -INFINITY;
    }
  }
}

// in-place row softmax, one query row per lane
__device__ __forceinline__ void softmax_phase(float* Ps, int S, int SP, int lane) {
  if (lane < S) {
    float* row = Ps + lane * SP;
    float m = -INFINITY;
    for (int j = 0; j < S; ++j) m = fmaxf(m, row[j]);
    float sum = 0.f;
    for (int j = 0; j < S; ++j) { const float e = __expf(row[j] - m); row[j] = e; sum += e; }
    const float inv = 1.f / sum;
    for (int j = 0; j < S; ++j) row[j] *= inv;
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
mha_fwd_small_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, T* __restrict__ out,
                     float* __restrict__ probs, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ float smem[];
  const int S = Tn + 1, SP = S | 1, F = H * D;
  constexpr int KP = AttSmem<D>::KP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Qs = smem + warp * AttSmem<D>::per_warp_fwd(S);
  float* Vs = Qs + S * D;
  float* Ks = Vs + S * D;
  float* Ps = Ks + S * KP;
  const float scale = rsqrtf((float)D);
  const long long total = (long long)B * H;
  for (long long bh = (long long)blockIdx.x * ATT_WARPS + warp; bh < total; bh += (long long)gridDim.x * ATT_WARPS) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    load_head<T, D>(base, 3 * F, Qs, D, S, lane);
    load_head<T, D>(base + F, 3 * F, Ks, KP, S, lane);
    load_head<T, D>(base + 2 * F, 3 * F, Vs, D, S, lane);
    const bool key_ok = lane < S && (lane == Tn || mask == nullptr || mask[(long long)b * Tn + lane] == 0);
    __syncwarp();
    scores_phase<D>(Qs, Ks, Ps, S, SP, lane, key_ok, scale);
    __syncwarp();
    softmax_phase(Ps, S, SP, lane);
    __syncwarp();
    if (probs != nullptr) {
      float* pg = probs + bh * S * S;
      for (int e = lane; e < S * S; e += 32) pg[e] = Ps[(e / S) * SP + (e % S)];
    }
    if (dc.thr) {
      for (int e = lane; e < S * S; e += 32) Ps[(e / S) * SP + (e % S)] *= drop1(dc, (uint64_t)(bh * S * S + e));
      __syncwarp();
    }
    // O = P V ; lane owns columns lane (+32).  Q is dead: reuse its storage for O.
    for (int i = 0; i < S; ++i) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < S; ++j) {
        const float p = Ps[i * SP + j];
        a0 = fmaf(p, Vs[j * D + lane], a0);
        if (D > 32) a1 = fmaf(p, Vs[j * D + lane + 32], a1);
      }
      Qs[i * D + lane] = a0;
      if (D > 32) Qs[i * D + lane + 32] = a1;
    }
    __syncwarp();
    store_head<T, D>(out + (long long)b * S * F + h * D, F, Qs, D, S, lane);
    __syncwarp();
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
mha_bwd_small_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, const T* __restrict__ dout,
                     T* __restrict__ dqkv, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ float smem[];
  const int S = Tn + 1, SP = S | 1, F = H * D;
  constexpr int KP = AttSmem<D>::KP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Qs = smem + warp * AttSmem<D>::per_warp_bwd(S);
  float* Vs = Qs + S * D;     // V, later the staging buffer for dQ / dK / dV
  float* dOs = Vs + S * D;
  float* Ks = dOs + S * D;    // padded rows (also holds V in padded form for the dP phase)
  float* Ps = Ks + S * KP;    // probabilities (then P*dropout)
  float* dSs = Ps + S * SP;   // dP then dS
  const float scale = rsqrtf((float)D);
  const long long total = (long long)B * H;
  for (long long bh = (long long)blockIdx.x * ATT_WARPS + warp; bh < total; bh += (long long)gridDim.x * ATT_WARPS) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    T* dbase = dqkv + (long long)b * S * 3 * F + h * D;
    load_head<T, D>(base, 3 * F, Qs, D, S, lane);
    load_head<T, D>(dout + (long long)b * S * F + h * D, F, dOs, D, S, lane);
    // dP[i][j] = dO_i . v_j : same shape as the score phase with V in the padded buffer
    load_head<T, D>(base + 2 * F, 3 * F, Ks, KP, S, lane);
    const bool key_ok = lane < S && (lane == Tn || mask == nullptr || mask[(long long)b * Tn + lane] == 0);
    __syncwarp();
    scores_phase<D>(dOs, Ks, dSs, S, SP, lane, true, 1.f);
    __syncwarp();
    load_head<T, D>(base + F, 3 * F, Ks, KP, S, lane);
    load_head<T, D>(base + 2 * F, 3 * F, Vs, D, S, lane);
    __syncwarp();
    scores_phase<D>(Qs, Ks, Ps, S, SP, lane, key_ok, scale);
    __syncwarp();
    softmax_phase(Ps, S, SP, lane);
    __syncwarp();
    // row i (one per lane): dP *= f ; dS = P * (dP - sum_j dP*P) ; P <- P*f (for dV)
    if (lane < S) {
      float* prow = Ps + lane * SP;
      float* drow = dSs + lane * SP;
      float dot = 0.f;
      for (int j = 0; j < S; ++j) {
        float f = 1.f;
        if (dc.thr) f = drop1(dc, (uint64_t)(bh * S * S + lane * S + j));
        const float dp = drow[j] * f;
        dot = fmaf(dp, prow[j], dot);
        drow[j] = dp;
      }
      for (int j = 0; j < S; ++j) {
        const float p = prow[j];
        drow[j] = p * (drow[j] - dot) * scale;  // fold the 1/sqrt(d) of the scores in here
        if (dc.thr) prow[j] = p * drop1(dc, (uint64_t)(bh * S * S + lane * S + j));
      }
    }
    __syncwarp();
    // dV[j][d] = sum_i Pd[i][j] dO[i][d]   (V itself is dead now: stage in Vs)
    for (int j = 0; j < S; ++j) {
      float a0 = 0.f, a1 = 0.f;
      for (int i = 0; i < S; ++i) {
        const float p = Ps[i * SP + j];
        a0 = fmaf(p, dOs[i * D + lane], a0);
        if (D > 32) a1 = fmaf(p, dOs[i * D + lane + 32], a1);
      }
      Vs[j * D + lane] = a0;
      if (D > 32) Vs[j * D + lane + 32] = a1;
    }
    __syncwarp();
    store_head<T, D>(dbase + 2 * F, 3 * F, Vs, D, S, lane);
    __syncwarp();
    // dQ[i][d] = sum_j dS[i][j] K[j][d]
    for (int i = 0; i < S; ++i) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < S; ++j) {
        const float ds = dSs[i * SP + j];
        a0 = fmaf(ds, Ks[j * KP + lane], a0);
        if (D > 32) a1 = fmaf(ds, Ks[j * KP + lane + 32], a1);
      }
      Vs[i * D + lane] = a0;
      if (D > 32) Vs[i * D + lane + 32] = a1;
    }
    __syncwarp();
    store_head<T, D>(dbase, 3 * F, Vs, D, S, lane);
    __syncwarp();
    // dK[j][d] = sum_i dS[i][j] Q[i][d]
    for (int j = 0; j < S; ++j) {
      float a0 = 0.f, a1 = 0.f;
      for (int i = 0; i < S; ++i) {
        const float ds = dSs[i * SP + j];
        a0 = fmaf(ds, Qs[i * D + lane], a0);
        if (D > 32) a1 = fmaf(ds, Qs[i * D + lane + 32], a1);
      }
      Vs[j * D + lane] = a0;
      if (D > 32) Vs[j * D + lane + 32] = a1;
    }
    __syncwarp();
    store_head<T, D>(dbase + F, 3 * F, Vs, D, S, lane);
    __syncwarp();
  }
}

template <typename T, int D>
static int mha_fwd_launch(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                          DropCfg dc, cudaStream_t st) {
  const int S = Tn + 1;
  const size_t smem = (size_t)ATT_WARPS * AttSmem<D>::per_warp_fwd(S) * sizeof(float);
  auto kern = mha_fwd_small_kernel<T, D>;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha_fwd)");
    configured = smem;
  }
  long long want = ((long long)B * H + ATT_WARPS - 1) / ATT_WARPS;
  long long cap = (long long)sm_count() * 8;
  kern<<<(unsigned)(want < cap ? want : cap), ATT_WARPS * 32, smem, st>>>((const T*)qkv, mask, (T*)out, probs, B, Tn, H, dc);
  MMER_LAUNCH_CHECK("mha_fwd_small_kernel");
  return 0;
}
template <typename T, int D>
static int mha_bwd_launch(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H,
                          DropCfg dc, cudaStream_t st) {
  const int S = Tn + 1;
  const size_t smem = (size_t)ATT_WARPS * AttSmem<D>::per_warp_bwd(S) * sizeof(float);
  auto kern = mha_bwd_small_kernel<T, D>;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha_bwd)");
    configured = smem;
  }
  long long want = ((long long)B * H + ATT_WARPS - 1) / ATT_WARPS;
  long long cap = (long long)sm_count() * 8;
  kern<<<(unsigned)(want < cap ? want : cap), ATT_WARPS * 32, smem, st>>>((const T*)qkv, mask, (const T*)dout, (T*)dqkv, B, Tn, H, dc);
  MMER_LAUNCH_CHECK("mha_bwd_small_kernel");
  return 0;
}

int mha_fwd_generic(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H,
                    int64_t d, int dtype, DropCfg dc, cudaStream_t st);
int mha_bwd_generic(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int64_t B, int64_t T,
                    int64_t H, int64_t d, int dtype, DropCfg dc, cudaStream_t st);

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_mha_fwd(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H,
                 int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  MMER_CHECK_ARG(qkv && out, "mha_fwd: null pointer");
  MMER_CHECK_ARG(d == 64 || d == 32, "mha_fwd: head dim %lld unsupported (32 or 64)", (long long)d);
  MMER_CHECK_ARG(T >= 1 && H >= 1, "mha_fwd: bad shape");
  if (B <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (T + 1 > 32) return mha_fwd_generic(qkv, mask, out, probs, B, T, H, d, dtype, dc, st);
  if (dtype == MMER_BF16)
    return d == 64 ? mha_fwd_launch<bf16, 64>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st)
                   : mha_fwd_launch<bf16, 32>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
  return d == 64 ? mha_fwd_launch<float, 64>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st)
                 : mha_fwd_launch<float, 32>(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
}

int mmer_mha_bwd(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int64_t B, int64_t T, int64_t H,
                 int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  MMER_CHECK_ARG(qkv && dout && dqkv, "mha_bwd: null pointer");
  MMER_CHECK_ARG(d == 64 || d == 32, "mha_bwd: head dim %lld unsupported (32 or 64)", (long long)d);
  MMER_CHECK_ARG(T >= 1 && H >= 1, "mha_bwd: bad shape");
  if (B <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (T + 1 > 32) return mha_bwd_generic(qkv, mask, dout, dqkv, B, T, H, d, dtype, dc, st);
  if (dtype == MMER_BF16)
    return d == 64 ? mha_bwd_launch<bf16, 64>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st)
                   : mha_bwd_launch<bf16, 32>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st);
  return d == 64 ? mha_bwd_launch<float, 64>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st)
                 : mha_bwd_launch<float, 32>(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st);
}

}  // extern "C"
