// Integrated Gradients around the fusion model: the two data-movement steps of
// captum.attr.IntegratedGradients.attribute as the reference calls it (train2.py:776-838,
// back-end/app/libs/inference.py:268-325), as single memory-bound kernels.  The model evaluations in between are the
// ordinary eval-mode forward/backward of the engine on the n_steps-times expanded batch.
//   expand : xs[k*B + b, :] = base[b, :] + alpha_k * (x[b, :] - base[b, :])          (step-major, like Captum's torch.cat)
//   reduce : attr[b, :]     = (x[b, :] - base[b, :]) * sum_k w_k * grad[k*B + b, :]   (multiply_by_inputs = True)
#include "common.cuh"

namespace mmer {

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
ig_expand_kernel(const TI* __restrict__ x, const TI* __restrict__ base, const float* __restrict__ alphas, TO* __restrict__ out,
                 long long n_per_step, int n_steps) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n_per_step) return;
  float xv[8], bv[8];
  load8(x + i, xv);
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = 0.f;
  if (base != nullptr) load8(base + i, bv);
  for (int k = blockIdx.y; k < n_steps; k += gridDim.y) {
    const float a = alphas[k];
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a, xv[j] - bv[j], bv[j]);
    store8(out + (long long)k * n_per_step + i, o);
  }
}

template <typename TI, typename TG>
__global__ void __launch_bounds__(256)
ig_reduce_kernel(const TG* __restrict__ grads, const TI* __restrict__ x, const TI* __restrict__ base,
                 const float* __restrict__ weights, float* __restrict__ attr, long long n_per_step, int n_steps) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n_per_step) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k = 0; k < n_steps; ++k) {
    const float w = weights[k];
    float g[8];
    load8(grads + (long long)k * n_per_step + i, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, g[j], acc[j]);
  }
  float xv[8], bv[8];
  load8(x + i, xv);
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = 0.f;
  if (base != nullptr) load8(base + i, bv);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= xv[j] - bv[j];
  store8(attr + i, acc);
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_ig_expand(const void* x, const void* base, const float* alphas, void* out, int64_t n_per_step, int64_t n_steps,
                   int in_dtype, int out_dtype, void* stream) {
  MMER_CHECK_ARG(x && alphas && out, "ig_expand: null pointer");
  MMER_CHECK_ARG(n_per_step > 0 && n_per_step % 8 == 0 && n_steps > 0, "ig_expand: sizes must be positive, elements %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((n_per_step / 8 + 255) / 256), (unsigned)(n_steps < 16 ? n_steps : 16));
  if (in_dtype == MMER_F32 && out_dtype == MMER_F32)
    ig_expand_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (const float*)base, alphas, (float*)out, n_per_step, (int)n_steps);
  else if (in_dtype == MMER_F32 && out_dtype == MMER_BF16)
    ig_expand_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, (const float*)base, alphas, (bf16*)out, n_per_step, (int)n_steps);
  else if (in_dtype == MMER_BF16 && out_dtype == MMER_BF16)
    ig_expand_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)base, alphas, (bf16*)out, n_per_step, (int)n_steps);
  else
    MMER_CHECK_ARG(false, "ig_expand: unsupported dtype pair");
  MMER_LAUNCH_CHECK("ig_expand_kernel");
  return 0;
}

int mmer_ig_reduce(const void* grads, const void* x, const void* base, const float* weights, float* attr, int64_t n_per_step,
                   int64_t n_steps, int in_dtype, int grad_dtype, void* stream) {
  MMER_CHECK_ARG(grads && x && weights && attr, "ig_reduce: null pointer");
  MMER_CHECK_ARG(n_per_step > 0 && n_per_step % 8 == 0 && n_steps > 0, "ig_reduce: sizes must be positive, elements %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_per_step / 8 + 255) / 256);
  if (in_dtype == MMER_F32 && grad_dtype == MMER_F32)
    ig_reduce_kernel<float, float><<<grid, 256, 0, st>>>((const float*)grads, (const float*)x, (const float*)base, weights, attr, n_per_step, (int)n_steps);
  else if (in_dtype == MMER_F32 && grad_dtype == MMER_BF16)
    ig_reduce_kernel<float, bf16><<<grid, 256, 0, st>>>((const bf16*)grads, (const float*)x, (const float*)base, weights, attr, n_per_step, (int)n_steps);
  else if (in_dtype == MMER_BF16 && grad_dtype == MMER_BF16)
    ig_reduce_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)grads, (const bf16*)x, (const bf16*)base, weights, attr, n_per_step, (int)n_steps);
  else
    MMER_CHECK_ARG(false, "ig_reduce: unsupported dtype pair");
  MMER_LAUNCH_CHECK("ig_reduce_kernel");
  return 0;
}

}  // extern "C"
