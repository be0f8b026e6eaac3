// Output layer Linear(hidden -> C, C <= 16) + softmax, its backward, and the fused
// FocalLoss / weighted cross-entropy forward+gradient kernel.  These touch B x C numbers:
// they are launch-latency bound, so each is a single small kernel.
#include "common.cuh"

namespace mmer {

static constexpr int MAXC = 16;

// one warp per sample: logits = h W^T + b ; probs = softmax(logits)
template <typename T>
__global__ void __launch_bounds__(256)
head_out_fwd_kernel(const T* __restrict__ h, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ logits, float* __restrict__ probs, int B, int K, int C) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  float acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float hv[8];
    load8(h + (long long)row * K + k, hv);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        float wv[8];
        load8(W + (long long)c * K + k, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c] = fmaf(hv[j], wv[j], acc[c]);
      }
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) { acc[c] = warp_sum(acc[c]) + bias[c]; m = fmaxf(m, acc[c]); }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c < C) sum += __expf(acc[c] - m);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        logits[(long long)row * C + c] = acc[c];
        if (probs != nullptr) probs[(long long)row * C + c] = __expf(acc[c] - m) / sum;
      }
    }
  }
}

// dh = dlogits W ; dW += dlogits^T h ; db += colsum(dlogits).  Warp w of a CTA owns the 256-column slab
// (w % nslab) of K and every (8/nslab)-th row of the CTA's row range: the lane's 8 columns of dW stay in registers
// for up to 8 classes at a time (8 x 8 accumulators); warps are combined through shared-memory atomics and each
// CTA adds its partial dW with one fp32 atomic per element, so the per-row traffic is the 16-byte read of h and
// the 16-byte write of dh.
static constexpr int HOB_WARPS = 8, HOB_CC = 8;
template <typename T>
__global__ void __launch_bounds__(HOB_WARPS * 32)
head_out_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ h, const float* __restrict__ W,
                    T* __restrict__ dh, float* __restrict__ dW, float* __restrict__ db, int B, int K, int C,
                    int rows_per_cta) {
  extern __shared__ float sW[];   // [min(C, 8)][K] partial dW of the current class group
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nslab = (K + 255) / 256;
  const int slab = warp % nslab, rgroup = warp / nslab, ngroups = HOB_WARPS / nslab;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(B, row0 + rows_per_cta);
  if (warp == 0 && lane < C) {
    float s = 0.f;
    for (int row = row0; row < row1; ++row) s += dlogits[(long long)row * C + lane];
    atomicAdd(db + lane, s);
  }
  const int k = slab * 256 + lane * 8;
  const bool active = rgroup < ngroups && k < K;
  for (int c0 = 0; c0 < C; c0 += HOB_CC) {
    const int cc = min(HOB_CC, C - c0);
    for (int i = threadIdx.x; i < cc * K; i += blockDim.x) sW[i] = 0.f;
    __syncthreads();
    if (active) {
      float acc[HOB_CC][8], wv[HOB_CC][8];
#pragma unroll
      for (int c = 0; c < HOB_CC; ++c) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[c][j] = 0.f; wv[c][j] = 0.f; }
        if (c < cc) load8(W + (long long)(c0 + c) * K + k, wv[c]);
      }
      for (int row = row0 + rgroup; row < row1; row += ngroups) {
        float hv[8], o[8], dl[HOB_CC];
        load8(h + (long long)row * K + k, hv);
#pragma unroll
        for (int c = 0; c < HOB_CC; ++c) dl[c] = (c < cc) ? __ldg(dlogits + (long long)row * C + c0 + c) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int c = 0; c < HOB_CC; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[c][j] = fmaf(dl[c], hv[j], acc[c][j]);
            o[j] = fmaf(dl[c], wv[c][j], o[j]);
          }
        if (dh != nullptr) {
          if (c0 > 0) {   // more than HOB_CC classes: add to what the earlier class groups wrote
            float prev[8];
            load8(dh + (long long)row * K + k, prev);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += prev[j];
          }
          store8(dh + (long long)row * K + k, o);
        }
      }
#pragma unroll
      for (int c = 0; c < HOB_CC; ++c)
        if (c < cc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(sW + c * K + k + j, acc[c][j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cc * K; i += blockDim.x) atomicAdd(dW + (long long)c0 * K + i, sW[i]);
    __syncthreads();
  }
}

// sum of class weights of the batch labels (denominator of the weighted CE mean)
__global__ void wce_den_kernel(const int64_t* __restrict__ labels, const float* __restrict__ w, float* den, int B, int C) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
    const int64_t y = labels[i];
    s += (y < 0 || y >= C) ? NAN : (w ? w[y] : 1.f);   // out-of-range label: poison the loss, never read out of bounds
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(den, s);
}

// one thread per sample
__global__ void __launch_bounds__(256)
loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ alpha,
            int kind, float gamma, int reduction, float* __restrict__ loss_out, float* __restrict__ per_sample,
            float* __restrict__ dlogits, const float* __restrict__ den, int B, int C, float grad_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lval = 0.f;
  if (i < B) {
    float x[MAXC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      x[c] = c < C ? logits[(long long)i * C + c] : -INFINITY;
      m = fmaxf(m, x[c]);
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) sum += c < C ? expf(x[c] - m) : 0.f;
    const float lse = m + logf(sum);
    const int64_t y64 = labels[i];
    const bool bad_label = y64 < 0 || y64 >= C;   // torch device-asserts here; we return a NaN loss and NaN gradients
    const int y = bad_label ? 0 : (int)y64;
    float xy = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) xy = (c == y) ? x[c] : xy;
    const float ce = lse - xy;
    const float w = bad_label ? NAN : (alpha ? alpha[y] : 1.f);
    float dce;  // d(per-sample loss)/d(ce), before the reduction factor
    if (kind == MMER_LOSS_FOCAL) {
      const float pt = expf(-ce);
      const float om = 1.f - pt;
      const float pw = powf(om, gamma);
      lval = w * pw * ce;
      // (1-pt)^g + g * pt * (1-pt)^(g-1) * ce ; guard (1-pt) == 0
      const float pw1 = om > 0.f ? powf(om, gamma - 1.f) : (gamma == 1.f ? 1.f : 0.f);
      dce = w * (pw + gamma * pt * pw1 * ce);
    } else {
      lval = w * ce;
      dce = w;
    }
    float red = 1.f;
    if (reduction == MMER_REDUCE_MEAN) red = (kind == MMER_LOSS_WCE) ? 1.f / *den : 1.f / (float)B;
    if (per_sample != nullptr) per_sample[i] = lval;
    if (dlogits != nullptr) {
      const float f = dce * red * grad_scale;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) dlogits[(long long)i * C + c] = f * (expf(x[c] - lse) - (c == y ? 1.f : 0.f));
    }
    lval *= red;
  }
  if (reduction != MMER_REDUCE_NONE && loss_out != nullptr) {
    __shared__ float sw[8];
    float s = warp_sum(lval);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sw[k];
      atomicAdd(loss_out, t);
    }
  }
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_head_out_fwd(const void* h, const float* W, const float* b, float* logits, float* probs, int64_t B,
                      int64_t K, int64_t C, int dtype, void* stream) {
  MMER_CHECK_ARG(h && W && b && logits, "head_out_fwd: null pointer");
  MMER_CHECK_ARG(C >= 1 && C <= MAXC && K % 8 == 0, "head_out_fwd: need C <= 16 and K %% 8 == 0 (C=%lld K=%lld)",
                 (long long)C, (long long)K);
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((B + 7) / 8);
  cudaError_t le;
  if (dtype == MMER_BF16)
    le = launch_dep(head_out_fwd_kernel<bf16>, dim3(grid), dim3(256), 0, st, 1, (const bf16*)h, W, b, logits, probs, (int)B, (int)K, (int)C);
  else
    le = launch_dep(head_out_fwd_kernel<float>, dim3(grid), dim3(256), 0, st, 1, (const float*)h, W, b, logits, probs, (int)B, (int)K, (int)C);
  if (le != cudaSuccess) return cuda_fail(le, "launch(head_out_fwd)");
  MMER_LAUNCH_CHECK("head_out_fwd_kernel");
  return 0;
}

int mmer_head_out_bwd(const float* dlogits, const void* h, const float* W, void* dh, float* dW, float* db, int64_t B,
                      int64_t K, int64_t C, int dtype, void* stream) {
  MMER_CHECK_ARG(dlogits && h && W && dW && db, "head_out_bwd: null pointer");
  MMER_CHECK_ARG(C >= 1 && C <= MAXC && K % 8 == 0, "head_out_bwd: need C <= 16 and K %% 8 == 0");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  MMER_CHECK_ARG(K <= 2048, "head_out_bwd: K above 2048 unsupported");
  const int nslab = (int)((K + 255) / 256);
  const int ngroups = HOB_WARPS / nslab;                 // rows in flight per CTA
  long long ctas = sm_count();
  long long rows_per_cta = (B + ctas - 1) / ctas;
  if (rows_per_cta < ngroups) rows_per_cta = ngroups;
  const unsigned grid = (unsigned)((B + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = (size_t)(C < HOB_CC ? C : HOB_CC) * K * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = dtype == MMER_BF16
                        ? cudaFuncSetAttribute(head_out_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(head_out_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(head_out_bwd)");
  }
  cudaError_t le;
  if (dtype == MMER_BF16)
    le = launch_dep(head_out_bwd_kernel<bf16>, dim3(grid), dim3(HOB_WARPS * 32), smem, st, 1, dlogits, (const bf16*)h, W, (bf16*)dh,
                    dW, db, (int)B, (int)K, (int)C, (int)rows_per_cta);
  else
    le = launch_dep(head_out_bwd_kernel<float>, dim3(grid), dim3(HOB_WARPS * 32), smem, st, 1, dlogits, (const float*)h, W,
                    (float*)dh, dW, db, (int)B, (int)K, (int)C, (int)rows_per_cta);
  if (le != cudaSuccess) return cuda_fail(le, "launch(head_out_bwd)");
  MMER_LAUNCH_CHECK("head_out_bwd_kernel");
  return 0;
}

int mmer_loss_fwd_bwd(const float* logits, const int64_t* labels, const float* alpha, int kind, float gamma,
                      int reduction, float* loss_out, float* per_sample, float* dlogits, float* scratch, int64_t B,
                      int64_t C, float grad_scale, void* stream) {
  MMER_CHECK_ARG(logits && labels, "loss: null pointer");
  MMER_CHECK_ARG(C >= 1 && C <= MAXC, "loss: C must be <= 16");
  MMER_CHECK_ARG(kind == MMER_LOSS_FOCAL || kind == MMER_LOSS_WCE, "loss: unknown kind %d", kind);
  MMER_CHECK_ARG(B > 0, "loss: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (loss_out != nullptr) {
    e = cudaMemsetAsync(loss_out, 0, sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset(loss)");
  }
  const float* den = nullptr;
  if (kind == MMER_LOSS_WCE && reduction == MMER_REDUCE_MEAN) {
    MMER_CHECK_ARG(scratch != nullptr, "loss: weighted CE needs scratch");
    e = cudaMemsetAsync(scratch, 0, sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset(den)");
    wce_den_kernel<<<(unsigned)((B + 255) / 256 < 64 ? (B + 255) / 256 : 64), 256, 0, st>>>(labels, alpha, scratch, (int)B, (int)C);
    MMER_LAUNCH_CHECK("wce_den_kernel");
    den = scratch;
  }
  loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(logits, labels, alpha, kind, gamma, reduction, loss_out,
                                                           per_sample, dlogits, den, (int)B, (int)C, grad_scale);
  MMER_LAUNCH_CHECK("loss_kernel");
  return 0;
}

}  // extern "C"
