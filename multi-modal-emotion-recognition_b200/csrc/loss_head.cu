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

// one warp per sample: dh = dlogits W ; dW += dlogits^T h ; db += colsum(dlogits)
template <typename T>
__global__ void __launch_bounds__(256)
head_out_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ h, const float* __restrict__ W,
                    T* __restrict__ dh, float* __restrict__ dW, float* __restrict__ db, int B, int K, int C) {
  extern __shared__ float sW[];  // [C][K] partial dW, then [C] partial db
  float* sdb = sW + (size_t)C * K;
  for (int i = threadIdx.x; i < C * K + C; i += blockDim.x) sW[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < B; row += gridDim.x * 8) {
    float dl[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) dl[c] = c < C ? dlogits[(long long)row * C + c] : 0.f;
    if (lane < C) atomicAdd(sdb + lane, dlogits[(long long)row * C + lane]);
    for (int k = lane * 8; k < K; k += 256) {
      float hv[8], o[8];
      load8(h + (long long)row * K + k, hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          float wv[8];
          load8(W + (long long)c * K + k, wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = fmaf(dl[c], wv[j], o[j]);
            atomicAdd(sW + c * K + k + j, dl[c] * hv[j]);
          }
        }
      }
      if (dh != nullptr) store8(dh + (long long)row * K + k, o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) atomicAdd(dW + i, sW[i]);
  if (threadIdx.x < C) atomicAdd(db + threadIdx.x, sdb[threadIdx.x]);
}

// sum of class weights of the batch labels (denominator of the weighted CE mean)
__global__ void wce_den_kernel(const int64_t* __restrict__ labels, const float* __restrict__ w, float* den, int B) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) s += w ? w[labels[i]] : 1.f;
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
    const int y = (int)labels[i];
    float xy = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) xy = (c == y) ? x[c] : xy;
    const float ce = lse - xy;
    const float w = alpha ? alpha[y] : 1.f;
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
  if (dtype == MMER_BF16) head_out_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)h, W, b, logits, probs, (int)B, (int)K, (int)C);
  else head_out_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)h, W, b, logits, probs, (int)B, (int)K, (int)C);
  MMER_LAUNCH_CHECK("head_out_fwd_kernel");
  return 0;
}

int mmer_head_out_bwd(const float* dlogits, const void* h, const float* W, void* dh, float* dW, float* db, int64_t B,
                      int64_t K, int64_t C, int dtype, void* stream) {
  MMER_CHECK_ARG(dlogits && h && W && dW && db, "head_out_bwd: null pointer");
  MMER_CHECK_ARG(C >= 1 && C <= MAXC && K % 8 == 0, "head_out_bwd: need C <= 16 and K %% 8 == 0");
  const size_t smem = ((size_t)C * K + C) * sizeof(float);
  MMER_CHECK_ARG(smem <= 48 * 1024, "head_out_bwd: C*K too large (%lld)", (long long)(C * K));
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  long long want = (B + 7) / 8;
  long long cap = sm_count();
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (dtype == MMER_BF16)
    head_out_bwd_kernel<bf16><<<grid, 256, smem, st>>>(dlogits, (const bf16*)h, W, (bf16*)dh, dW, db, (int)B, (int)K, (int)C);
  else
    head_out_bwd_kernel<float><<<grid, 256, smem, st>>>(dlogits, (const float*)h, W, (float*)dh, dW, db, (int)B, (int)K, (int)C);
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
    wce_den_kernel<<<(unsigned)((B + 255) / 256 < 64 ? (B + 255) / 256 : 64), 256, 0, st>>>(labels, alpha, scratch, (int)B);
    MMER_LAUNCH_CHECK("wce_den_kernel");
    den = scratch;
  }
  loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(logits, labels, alpha, kind, gamma, reduction, loss_out,
                                                           per_sample, dlogits, den, (int)B, (int)C, grad_scale);
  MMER_LAUNCH_CHECK("loss_kernel");
  return 0;
}

}  // extern "C"
