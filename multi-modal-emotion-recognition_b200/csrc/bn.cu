// BatchNorm1d over the rows of an [N, C] matrix, for the train.py variant of the model
// (bn_video / bn_audio over all B*T resp. B projected rows, padded rows included, and
// bn_fc1 in the head).  Column statistics are reduced with the same 32-lane x 8-column
// coalesced access pattern as the other row kernels, then applied in one elementwise pass.
#include "common.cuh"

namespace mmer {

static constexpr float BN_EPS = 1e-5f;

// mode 0: s0 += x ; mode 1: s0 += (x-mean)^2 ; mode 2: s0 += dy_eff, s1 += dy_eff * xhat
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ stats,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ s0,
                 float* __restrict__ s1, long long N, int C, int relu, DropCfg dc) {
  __shared__ float sred[2][8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float a0[8], a1[8], mean[8], rstd[8], g[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; mean[j] = 0.f; rstd[j] = 1.f; g[j] = 1.f; be[j] = 0.f; }
  if (c < C) {
    if (MODE >= 1) load8(stats + c, mean);
    if (MODE == 2) { load8(stats + C + c, rstd); if (relu) { load8(gamma + c, g); load8(beta + c, be); } }
    for (long long r = (long long)blockIdx.y * 8 + warp; r < N; r += (long long)gridDim.y * 8) {
      float v[8];
      load8(x + r * C + c, v);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a0[j] += v[j];
      } else if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean[j]; a0[j] += d * d; }
      } else {
        float d[8];
        load8(dy + r * C + c, d);
        if (dc.thr) {
          float f[8];
          drop8(dc, (uint64_t)(r * C + c), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[j] - mean[j]) * rstd[j];
          if (relu && !(xh * g[j] + be[j] > 0.f)) d[j] = 0.f;
          a0[j] += d[j];
          a1[j] += d[j] * xh;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sred[0][warp][lane * 8 + j] = a0[j]; sred[1][warp][lane * 8 + j] = a1[j]; }
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t0 += sred[0][w][threadIdx.x]; t1 += sred[1][w][threadIdx.x]; }
    atomicAdd(s0 + cc, t0);
    if (MODE == 2) atomicAdd(s1 + cc, t1);
  }
}

// phase: 0 -> stats[0..C) = sum/N (mean) ; 1 -> stats[C..2C) = rstd from sum of squared deviations,
// and the running statistics update (momentum, unbiased variance) ; 2 -> eval: stats from running
__global__ void bn_finalize_kernel(float* __restrict__ stats, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long N, int C, int phase, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (phase == 0) {
    stats[c] = stats[c] / (float)N;
  } else if (phase == 1) {
    const float var = stats[C + c] / (float)N;
    stats[C + c] = rsqrtf(var + BN_EPS);
    if (running_mean != nullptr) {
      const float unb = N > 1 ? var * ((float)N / (float)(N - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * stats[c];
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  } else {
    stats[c] = running_mean[c];
    stats[C + c] = rsqrtf(running_var[c] + BN_EPS);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, T* __restrict__ y, long long N, int C, int relu, DropCfg dc) {
  const long long total = N * C / 8;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long off = e * 8;
    const int c = (int)(off % C);
    float v[8], mean[8], rstd[8], g[8], be[8], o[8];
    load8(x + off, v);
    load8(stats + c, mean);
    load8(stats + C + c, rstd);
    load8(gamma + c, g);
    load8(beta + c, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = (v[j] - mean[j]) * rstd[j] * g[j] + be[j];
      if (relu) o[j] = fmaxf(o[j], 0.f);
    }
    if (dc.thr) {
      float f[8];
      drop8(dc, (uint64_t)off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= f[j];
    }
    store8(y + off, o);
  }
}

// dx = g*rstd*(dy - sum(dy)/N - xhat*sum(dy*xhat)/N) in training; g*rstd*dy in eval.
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ sums,
                    T* __restrict__ dx, long long N, long long Nstat, int C, int relu, int training, DropCfg dc) {
  const long long total = N * C / 8;
  const float invN = 1.f / (float)Nstat;   // rows the statistics were taken over (the global batch under SyncBatchNorm)
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long off = e * 8;
    const int c = (int)(off % C);
    float d[8], v[8], mean[8], rstd[8], g[8], be[8], s0[8], s1[8], o[8];
    load8(dy + off, d);
    load8(x + off, v);
    load8(stats + c, mean);
    load8(stats + C + c, rstd);
    load8(gamma + c, g);
    load8(beta + c, be);
    load8(sums + c, s0);
    load8(sums + C + c, s1);
    if (dc.thr) {
      float f[8];
      drop8(dc, (uint64_t)off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] *= f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (v[j] - mean[j]) * rstd[j];
      if (relu && !(xh * g[j] + be[j] > 0.f)) d[j] = 0.f;
      o[j] = training ? g[j] * rstd[j] * (d[j] - s0[j] * invN - xh * s1[j] * invN) : g[j] * rstd[j] * d[j];
    }
    store8(dx + off, o);
  }
}

__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta != nullptr) dbeta[c] += sums[c];
  if (dgamma != nullptr) dgamma[c] += sums[C + c];
}

static dim3 reduce_grid(long long N, long long C) {
  int gx = (int)((C + 255) / 256);
  int gy = (sm_count() * 4) / gx;
  long long maxy = (N + 7) / 8;
  if (gy > maxy) gy = (int)maxy;
  if (gy < 1) gy = 1;
  return dim3(gx, gy);
}
static unsigned ew_grid(long long N, long long C) {
  long long want = (N * C / 8 + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

// SyncBatchNorm (mmer_model.bn_sync): all-reduce a buffer of partial column sums over the replicas, on the stream
static int bn_sync_call(const BnSync* sy, float* buf, long long n, cudaStream_t st) {
  if (sy == nullptr || sy->fn == nullptr || sy->world <= 1) return 0;
  const int rc = sy->fn(sy->user, buf, (int64_t)n, (void*)st);
  if (rc != 0) { set_error("bn_sync callback failed (%d)", rc); return MMER_ERR_ARG; }
  return 0;
}
static long long bn_rows(const BnSync* sy, long long N) {
  return (sy != nullptr && sy->fn != nullptr && sy->world > 1) ? N * sy->world : N;
}

template <typename T>
static int bn_fwd_t(const void* x, const float* gamma, const float* beta, float* rm, float* rv, void* y, float* stats,
                    long long N, long long C, int training, int relu, float momentum, DropCfg dc, cudaStream_t st,
                    const BnSync* sy) {
  const unsigned cb = (unsigned)((C + 127) / 128);
  if (training) {
    const long long Ng = bn_rows(sy, N);   // rows of the whole (global) batch
    cudaError_t e = cudaMemsetAsync(stats, 0, 2 * C * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset(bn stats)");
    bn_reduce_kernel<T, 0><<<reduce_grid(N, C), 256, 0, st>>>((const T*)x, nullptr, stats, nullptr, nullptr, stats, nullptr, N, (int)C, 0, dc);
    MMER_TRY(bn_sync_call(sy, stats, C, st));
    bn_finalize_kernel<<<cb, 128, 0, st>>>(stats, rm, rv, Ng, (int)C, 0, momentum);
    bn_reduce_kernel<T, 1><<<reduce_grid(N, C), 256, 0, st>>>((const T*)x, nullptr, stats, nullptr, nullptr, stats + C, nullptr, N, (int)C, 0, dc);
    MMER_TRY(bn_sync_call(sy, stats + C, C, st));
    bn_finalize_kernel<<<cb, 128, 0, st>>>(stats, rm, rv, Ng, (int)C, 1, momentum);
  } else {
    bn_finalize_kernel<<<cb, 128, 0, st>>>(stats, rm, rv, N, (int)C, 2, momentum);
  }
  bn_apply_kernel<T><<<ew_grid(N, C), 256, 0, st>>>((const T*)x, stats, gamma, beta, (T*)y, N, (int)C, relu, dc);
  MMER_LAUNCH_CHECK("bn_fwd");
  count_launch(training ? 4 : 1);
  return 0;
}

template <typename T>
static int bn_bwd_t(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, void* dx,
                    float* dgamma, float* dbeta, float* scratch, long long N, long long C, int relu, int training,
                    DropCfg dc, cudaStream_t st, const BnSync* sy) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * C * sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "memset(bn scratch)");
  bn_reduce_kernel<T, 2><<<reduce_grid(N, C), 256, 0, st>>>((const T*)x, (const T*)dy, stats, gamma, beta, scratch, scratch + C, N, (int)C, relu, dc);
  // the parameter gradients take this replica's LOCAL sums (the gradient exchange adds the replicas later); the input
  // gradient needs the sums over the whole batch
  bn_param_grad_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(scratch, dgamma, dbeta, (int)C);
  if (training) MMER_TRY(bn_sync_call(sy, scratch, 2 * C, st));
  bn_bwd_apply_kernel<T><<<ew_grid(N, C), 256, 0, st>>>((const T*)dy, (const T*)x, stats, gamma, beta, scratch, (T*)dx, N, training ? bn_rows(sy, N) : N, (int)C, relu, training, dc);
  MMER_LAUNCH_CHECK("bn_bwd");
  count_launch(2);
  return 0;
}

int bn_fwd_sync(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var, void* y,
                float* stats_out, int64_t N, int64_t C, int dtype, int training, int relu, float momentum, float drop_p,
                uint64_t seed, uint32_t site, cudaStream_t st, const BnSync* sy) {
  MMER_CHECK_ARG(x && gamma && beta && y && stats_out, "bn_fwd: null pointer");
  MMER_CHECK_ARG(C > 0 && C % 8 == 0, "bn_fwd: C must be a multiple of 8");
  MMER_CHECK_ARG(training || (running_mean && running_var), "bn_fwd: eval mode needs running statistics");
  if (N <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  return dtype == MMER_BF16
             ? bn_fwd_t<bf16>(x, gamma, beta, running_mean, running_var, y, stats_out, N, C, training, relu, momentum, dc, st, sy)
             : bn_fwd_t<float>(x, gamma, beta, running_mean, running_var, y, stats_out, N, C, training, relu, momentum, dc, st, sy);
}

int bn_bwd_sync(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, void* dx,
                float* dgamma, float* dbeta, float* scratch, int64_t N, int64_t C, int dtype, int training, int relu,
                float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, const BnSync* sy) {
  MMER_CHECK_ARG(dy && x && stats && gamma && beta && dx && scratch, "bn_bwd: null pointer");
  MMER_CHECK_ARG(C > 0 && C % 8 == 0, "bn_bwd: C must be a multiple of 8");
  if (N <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  return dtype == MMER_BF16
             ? bn_bwd_t<bf16>(dy, x, stats, gamma, beta, dx, dgamma, dbeta, scratch, N, C, relu, training, dc, st, sy)
             : bn_bwd_t<float>(dy, x, stats, gamma, beta, dx, dgamma, dbeta, scratch, N, C, relu, training, dc, st, sy);
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_bn_fwd(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var, void* y,
                float* stats_out, int64_t N, int64_t C, int dtype, int training, int relu, float momentum,
                float drop_p, uint64_t seed, uint32_t site, void* stream) {
  return bn_fwd_sync(x, gamma, beta, running_mean, running_var, y, stats_out, N, C, dtype, training, relu, momentum, drop_p,
                     seed, site, (cudaStream_t)stream, nullptr);
}

int mmer_bn_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, void* dx,
                float* dgamma, float* dbeta, float* scratch, int64_t N, int64_t C, int dtype, int training, int relu,
                float drop_p, uint64_t seed, uint32_t site, void* stream) {
  return bn_bwd_sync(dy, x, stats, gamma, beta, dx, dgamma, dbeta, scratch, N, C, dtype, training, relu, drop_p, seed, site,
                     (cudaStream_t)stream, nullptr);
}

}  // extern "C"
