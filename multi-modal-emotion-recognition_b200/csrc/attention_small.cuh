// Multi-head self-attention over the short token sequence [T video tokens ; audio token].
//
// Short-sequence path (S = T+1 <= 32, the benchmark shape is S = 17): one warp owns one
// (sample, head) at a time.  The work is register-tiled so that shared memory only carries the
// operands that must be broadcast:
//   scores   lane j keeps key row j in registers (read straight from the packed in_proj output,
//            128 B per row); query rows are broadcast from shared memory as float4
//   softmax  one query row per lane, in shared memory, masked keys get probability exactly 0
//   P.V      lane l keeps columns (2l, 2l+1) of every value row in registers; probability rows are
//            broadcast as float4; results leave as coalesced 128 B rows
// The backward kernel recomputes the probabilities (nothing but the packed QKV tensor is kept from
// the forward pass) and produces dQ, dK, dV with the same two patterns.  SP is S rounded up to a
// multiple of 4 and is a template parameter so that the per-lane register tiles are fully unrolled.
#pragma once
#include "common.cuh"

namespace mmer {

static constexpr int ATT_WARPS = 4;

// cooperative load of a [S][D] head slice (row stride ld elements) into fp32 smem (row stride D)
template <typename T, int D>
__device__ __forceinline__ void load_head(const T* __restrict__ g, long long ld, float* s, int S, int lane) {
  constexpr int LPR = D / 8;     // lanes per row
  constexpr int RPP = 32 / LPR;  // rows per pass
  const int c = (lane % LPR) * 8;
  for (int r0 = 0; r0 < S; r0 += RPP) {
    const int r = r0 + lane / LPR;
    if (r < S) {
      float v[8];
      load8(g + (long long)r * ld + c, v);
      *reinterpret_cast<float4*>(s + r * D + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(s + r * D + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// one row of D elements -> registers (zeros when !ok)
template <typename T, int D>
__device__ __forceinline__ void load_row(const T* __restrict__ g, bool ok, float (&r)[D]) {
#pragma unroll
  for (int c = 0; c < D / 8; ++c) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ok) load8(g + c * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[c * 8 + i] = v[i];
  }
}

// CPL = D/32 consecutive columns per lane
template <int CPL> struct ColVec;
template <> struct ColVec<2> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[2]) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[2]) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
  }
};
template <> struct ColVec<1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[1]) { *p = __float2bfloat16_rn(v[0]); }
};

// this lane's CPL columns of rows 0..SP-1 of a [S][*] matrix (row stride ld), zeros for rows >= S
template <typename T, int SP, int CPL>
__device__ __forceinline__ void load_cols(const T* __restrict__ g, long long ld, int S, float (&c)[SP][CPL]) {
#pragma unroll
  for (int j = 0; j < SP; ++j) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) c[j][k] = 0.f;
    if (j < S) ColVec<CPL>::load(g + (long long)j * ld, c[j]);
  }
}

// out[r][cols] = sum_j W[r][j] * c[j][cols] for r < S; W rows (SP floats, 16-byte aligned) broadcast from smem
template <typename T, int SP, int CPL>
__device__ __forceinline__ void rows_times_cols(const float* __restrict__ W, const float (&c)[SP][CPL],
                                                T* __restrict__ out, long long ld, int S) {
#pragma unroll
  for (int r = 0; r < SP; ++r) {
    if (r < S) {
      float acc[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < SP / 4; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W + r * SP + j4 * 4);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          acc[k] = fmaf(w.x, c[j4 * 4 + 0][k], acc[k]);
          acc[k] = fmaf(w.y, c[j4 * 4 + 1][k], acc[k]);
          acc[k] = fmaf(w.z, c[j4 * 4 + 2][k], acc[k]);
          acc[k] = fmaf(w.w, c[j4 * 4 + 3][k], acc[k]);
        }
      }
      ColVec<CPL>::store(out + (long long)r * ld, acc);
    }
  }
}

// Out[i][lane] = scale * (X_i . kreg) for i < S, rows X_i broadcast from smem; masked lanes get `fill`
template <int D, int SP>
__device__ __forceinline__ void dots_phase(const float* __restrict__ Xs, const float (&kreg)[D], float* __restrict__ Out,
                                           int S, int lane, bool ok, float scale, float fill) {
  for (int i = 0; i < S; ++i) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int d = 0; d < D / 4; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(Xs + i * D + d * 4);
      a0 = fmaf(q.x, kreg[d * 4 + 0], a0);
      a1 = fmaf(q.y, kreg[d * 4 + 1], a1);
      a2 = fmaf(q.z, kreg[d * 4 + 2], a2);
      a3 = fmaf(q.w, kreg[d * 4 + 3], a3);
    }
    if (lane < SP) Out[i * SP + lane] = ok ? ((a0 + a1) + (a2 + a3)) * scale : fill;
  }
}

template <int D, int SP> struct AttSmem {
  static constexpr int FWD = SP * D + SP * SP;           // Q rows, P
  static constexpr int BWD = 2 * SP * D + 4 * SP * SP;   // Q rows, dO rows, P, dS, Pd^T, dS^T
};

template <typename T, int D, int SP>
__global__ void __launch_bounds__(ATT_WARPS * 32)
mha_fwd_small_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, T* __restrict__ out,
                     float* __restrict__ probs, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CPL = D / 32;
  const int S = Tn + 1, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Qs = smem + warp * AttSmem<D, SP>::FWD;
  float* Ps = Qs + SP * D;
  const float scale = rsqrtf((float)D);
  const long long total = (long long)B * H;
  for (long long bh = (long long)blockIdx.x * ATT_WARPS + warp; bh < total; bh += (long long)gridDim.x * ATT_WARPS) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    float kreg[D];
    load_row<T, D>(base + F + (long long)lane * 3 * F, lane < S, kreg);
    load_head<T, D>(base, 3 * F, Qs, S, lane);
    const bool key_ok = lane < S && (lane == Tn || mask == nullptr || mask[(long long)b * Tn + lane] == 0);
    __syncwarp();
    dots_phase<D, SP>(Qs, kreg, Ps, S, lane, key_ok, scale, -INFINITY);
    float vcol[SP][CPL];
    load_cols<T, SP, CPL>(base + 2 * F + lane * CPL, 3 * F, S, vcol);
    __syncwarp();
    // softmax of row `lane`; columns >= S of the padded row end up exactly 0
    if (lane < S) {
      float* row = Ps + lane * SP;
      float m = -INFINITY;
      for (int j = 0; j < S; ++j) m = fmaxf(m, row[j]);
      float sum = 0.f;
      for (int j = 0; j < S; ++j) { const float e = __expf(row[j] - m); row[j] = e; sum += e; }
      const float inv = 1.f / sum;
      float* pg = probs ? probs + bh * S * S + (long long)lane * S : nullptr;
      for (int j = 0; j < S; ++j) {
        float p = row[j] * inv;
        if (pg) pg[j] = p;
        if (dc.thr) p *= drop1(dc, att_drop_index(bh * S + lane, j, att_drop_stride(S)));
        row[j] = p;
      }
      for (int j = S; j < SP; ++j) row[j] = 0.f;
    }
    __syncwarp();
    rows_times_cols<T, SP, CPL>(Ps, vcol, out + (long long)b * S * F + h * D + lane * CPL, F, S);
    __syncwarp();
  }
}

template <typename T, int D, int SP>
__global__ void __launch_bounds__(ATT_WARPS * 32)
mha_bwd_small_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask, const T* __restrict__ dout,
                     T* __restrict__ dqkv, int B, int Tn, int H, DropCfg dc) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CPL = D / 32;
  const int S = Tn + 1, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Qs = smem + warp * AttSmem<D, SP>::BWD;
  float* dOs = Qs + SP * D;
  float* Ps = dOs + SP * D;    // scores, then probabilities            [i][j]
  float* dSs = Ps + SP * SP;   // dP, then dS                            [i][j]
  float* PdT = dSs + SP * SP;  // probabilities after dropout, transposed [j][i]
  float* dST = PdT + SP * SP;  // dS transposed                          [j][i]
  const float scale = rsqrtf((float)D);
  const long long total = (long long)B * H;
  for (long long bh = (long long)blockIdx.x * ATT_WARPS + warp; bh < total; bh += (long long)gridDim.x * ATT_WARPS) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    const T* base = qkv + (long long)b * S * 3 * F + h * D;
    const T* dobase = dout + (long long)b * S * F + h * D;
    T* dbase = dqkv + (long long)b * S * 3 * F + h * D;
    const bool key_ok = lane < S && (lane == Tn || mask == nullptr || mask[(long long)b * Tn + lane] == 0);
    {
      float vreg[D];
      load_row<T, D>(base + 2 * F + (long long)lane * 3 * F, lane < S, vreg);
      load_head<T, D>(dobase, F, dOs, S, lane);
      load_head<T, D>(base, 3 * F, Qs, S, lane);
      __syncwarp();
      dots_phase<D, SP>(dOs, vreg, dSs, S, lane, true, 1.f, 0.f);   // dP[i][j] = dO_i . v_j
    }
    {
      float kreg[D];
      load_row<T, D>(base + F + (long long)lane * 3 * F, lane < S, kreg);
      dots_phase<D, SP>(Qs, kreg, Ps, S, lane, key_ok, scale, -INFINITY);
    }
    __syncwarp();
    // row i = lane: P = softmax(S_i); dP *= dropout; dS = P * (dP - sum_j dP*P) * scale; Pd = P * dropout
    if (lane < SP) {
      if (lane < S) {
        float* prow = Ps + lane * SP;
        float* drow = dSs + lane * SP;
        float m = -INFINITY;
        for (int j = 0; j < S; ++j) m = fmaxf(m, prow[j]);
        float sum = 0.f;
        for (int j = 0; j < S; ++j) { const float e = __expf(prow[j] - m); prow[j] = e; sum += e; }
        const float inv = 1.f / sum;
        float dot = 0.f;
        for (int j = 0; j < S; ++j) {
          const float p = prow[j] * inv;
          float f = 1.f;
          if (dc.thr) f = drop1(dc, att_drop_index(bh * S + lane, j, att_drop_stride(S)));
          const float dp = drow[j] * f;
          dot = fmaf(dp, p, dot);
          drow[j] = dp;
          prow[j] = p;
          PdT[j * SP + lane] = p * f;
        }
        for (int j = 0; j < S; ++j) {
          const float ds = prow[j] * (drow[j] - dot) * scale;  // the 1/sqrt(d) of the scores is folded in here
          drow[j] = ds;
          dST[j * SP + lane] = ds;
        }
        for (int j = S; j < SP; ++j) { drow[j] = 0.f; PdT[j * SP + lane] = 0.f; dST[j * SP + lane] = 0.f; }
      } else {
        for (int j = 0; j < SP; ++j) { dSs[lane * SP + j] = 0.f; PdT[j * SP + lane] = 0.f; dST[j * SP + lane] = 0.f; }
      }
    }
    __syncwarp();
    float col[SP][CPL];
    // dV[j] = sum_i Pd[i][j] dO[i]
    load_cols<T, SP, CPL>(dobase + lane * CPL, F, S, col);
    rows_times_cols<T, SP, CPL>(PdT, col, dbase + 2 * F + lane * CPL, 3 * F, S);
    // dQ[i] = sum_j dS[i][j] K[j]
    load_cols<T, SP, CPL>(base + F + lane * CPL, 3 * F, S, col);
    rows_times_cols<T, SP, CPL>(dSs, col, dbase + lane * CPL, 3 * F, S);
    // dK[j] = sum_i dS[i][j] Q[i]
    load_cols<T, SP, CPL>(base + lane * CPL, 3 * F, S, col);
    rows_times_cols<T, SP, CPL>(dST, col, dbase + F + lane * CPL, 3 * F, S);
    __syncwarp();
  }
}

template <typename K>
static int att_configure(K kern, size_t smem, size_t* configured, int* blocks_per_sm) {
  if (smem > *configured || *blocks_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha)");
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, ATT_WARPS * 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy(mha)");
    *blocks_per_sm = n > 0 ? n : 1;
    *configured = smem;
  }
  return 0;
}

template <typename T, int D, int SP>
static int mha_fwd_launch(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                          DropCfg dc, cudaStream_t st) {
  const size_t smem = (size_t)ATT_WARPS * AttSmem<D, SP>::FWD * sizeof(float);
  auto kern = mha_fwd_small_kernel<T, D, SP>;
  static size_t configured = 0;
  static int bps = 0;
  MMER_TRY(att_configure(kern, smem, &configured, &bps));
  const long long want = ((long long)B * H + ATT_WARPS - 1) / ATT_WARPS;
  const long long cap = (long long)sm_count() * bps;
  kern<<<(unsigned)(want < cap ? want : cap), ATT_WARPS * 32, smem, st>>>((const T*)qkv, mask, (T*)out, probs, B, Tn, H, dc);
  MMER_LAUNCH_CHECK("mha_fwd_small_kernel");
  return 0;
}
template <typename T, int D, int SP>
static int mha_bwd_launch(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H,
                          DropCfg dc, cudaStream_t st) {
  const size_t smem = (size_t)ATT_WARPS * AttSmem<D, SP>::BWD * sizeof(float);
  auto kern = mha_bwd_small_kernel<T, D, SP>;
  static size_t configured = 0;
  static int bps = 0;
  MMER_TRY(att_configure(kern, smem, &configured, &bps));
  const long long want = ((long long)B * H + ATT_WARPS - 1) / ATT_WARPS;
  const long long cap = (long long)sm_count() * bps;
  kern<<<(unsigned)(want < cap ? want : cap), ATT_WARPS * 32, smem, st>>>((const T*)qkv, mask, (const T*)dout, (T*)dqkv, B, Tn,
                                                                          H, dc);
  MMER_LAUNCH_CHECK("mha_bwd_small_kernel");
  return 0;
}

template <typename T, int D>
static int mha_fwd_sp(int SP, const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                      DropCfg dc, cudaStream_t st) {
  switch (SP) {
    case 4: return mha_fwd_launch<T, D, 4>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 8: return mha_fwd_launch<T, D, 8>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 12: return mha_fwd_launch<T, D, 12>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 16: return mha_fwd_launch<T, D, 16>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 20: return mha_fwd_launch<T, D, 20>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 24: return mha_fwd_launch<T, D, 24>(qkv, mask, out, probs, B, Tn, H, dc, st);
    case 28: return mha_fwd_launch<T, D, 28>(qkv, mask, out, probs, B, Tn, H, dc, st);
    default: return mha_fwd_launch<T, D, 32>(qkv, mask, out, probs, B, Tn, H, dc, st);
  }
}
template <typename T, int D>
static int mha_bwd_sp(int SP, const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H,
                      DropCfg dc, cudaStream_t st) {
  switch (SP) {
    case 4: return mha_bwd_launch<T, D, 4>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 8: return mha_bwd_launch<T, D, 8>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 12: return mha_bwd_launch<T, D, 12>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 16: return mha_bwd_launch<T, D, 16>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 20: return mha_bwd_launch<T, D, 20>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 24: return mha_bwd_launch<T, D, 24>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    case 28: return mha_bwd_launch<T, D, 28>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
    default: return mha_bwd_launch<T, D, 32>(qkv, mask, dout, dqkv, B, Tn, H, dc, st);
  }
}

}  // namespace mmer
