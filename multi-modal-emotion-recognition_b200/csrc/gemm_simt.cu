// fp32 FMA GEMM with the same operand-major / epilogue contract as the tcgen05 kernel.
// This is the fp32 parity mode (logits within 1e-4 of the reference need true fp32
// products; tcgen05 offers tf32 at best).  64x64 tile, 16-deep k slices, 4x4 per thread.
#include "common.cuh"

namespace mmer {

static constexpr int TS = 64, TK = 16;

struct SimtParams {
  const float* A; const float* B; float* D;
  const float* bias; const float* residual; const float* gate;
  long long M, N, K, sam, sak, sbn, sbk, ldd;
  int accumulate, relu;
  float gate_scale;
  DropCfg drop;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[TK][TS + 4];
  __shared__ float Bs[TK][TS + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.y * TS, n0 = (long long)blockIdx.x * TS;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: pick the one that walks the contiguous dimension with consecutive threads
  const bool a_kfast = p.sak == 1, b_kfast = p.sbk == 1;
  for (long long k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + r * 256;  // 0..1023
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const long long gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < p.K) ? p.A[gm * p.sam + gk * p.sak] : 0.f;
      int nn;
      if (b_kfast) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const long long gn = n0 + nn, gk2 = k0 + kk;
      Bs[kk][nn] = (gn < p.N && gk2 < p.K) ? p.B[gn * p.sbn + gk2 * p.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      const long long off = m * p.ldd + n;
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.drop.thr) v *= drop1(p.drop, (uint64_t)off);
      if (p.gate) v *= (p.gate[off] > 0.f) ? p.gate_scale : 0.f;
      if (p.residual) v += p.residual[off];
      if (p.accumulate) v += p.D[off];
      p.D[off] = v;
    }
  }
}

int gemm_simt(const mmer_gemm_args& a, cudaStream_t st) {
  MMER_CHECK_ARG(a.in_dtype == MMER_F32 && a.out_dtype == MMER_F32, "gemm_simt: fp32 in/out only");
  MMER_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm_simt: empty problem");
  MMER_CHECK_ARG(a.relu_mask_out == nullptr && a.gate_bits == nullptr, "gemm_simt: bit masks are a bf16 (tcgen05) feature");
  SimtParams p;
  p.A = (const float*)a.A; p.B = (const float*)a.B; p.D = (float*)a.D;
  p.bias = a.bias; p.residual = (const float*)a.residual; p.gate = (const float*)a.gate;
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldd = a.ldd;
  if (a.a_major == MMER_MAJOR_K) { p.sam = a.lda; p.sak = 1; } else { p.sam = 1; p.sak = a.lda; }
  if (a.b_major == MMER_MAJOR_K) { p.sbn = a.ldb; p.sbk = 1; } else { p.sbn = 1; p.sbk = a.ldb; }
  p.accumulate = a.accumulate; p.relu = a.relu; p.gate_scale = a.gate_scale;
  p.drop = make_drop(a.drop_p, a.seed, a.drop_site);
  dim3 grid(ceil_div(a.N, TS), ceil_div(a.M, TS));
  gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
  MMER_LAUNCH_CHECK("gemm_simt_kernel");
  return 0;
}

}  // namespace mmer
