// Batch assembly on the device: the step right before the fusion model in the reference's loop.
//   feature statistics : train2.py:430-447  (global mean / unbiased std + 1e-6 over all frames, per feature)
//   collate            : train2.py:418-440 of the closure collate_fn -> pad_sequence(videos), stack(audios), labels,
//                        pad_sequence(masks, padding_value=True); with the z-score of train2.py:443-447 applied on the fly
// The whole feature set stays resident in HBM as one ragged [total_frames, Dv] array plus frame offsets; a batch is ONE
// gather kernel that reads each needed frame once and writes the padded, normalised, (optionally bf16) batch and its
// padding mask: pure byte movement, HBM-bound.
#include "common.cuh"

namespace mmer {

// column sums of x[R, D] (pass = 0) or of (x - mean)^2 (pass = 1) in double: thread = column, CTA = row stripe
__global__ void __launch_bounds__(256)
feature_moment_kernel(const float* __restrict__ x, long long R, int D, const double* __restrict__ mean_sum, int pass,
                      double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const double mu = pass ? mean_sum[c] / (double)R : 0.0;
  double acc = 0.0;
  for (long long r = blockIdx.y; r < R; r += gridDim.y) {
    const double v = (double)x[r * D + c] - mu;
    acc += pass ? v * v : v;
  }
  atomicAdd(out + c, acc);
}
// mean = sum / R;  std = sqrt(ss / (R - 1)) + eps   (torch.std default: Bessel's correction; R = 1 gives NaN like torch)
__global__ void feature_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ ss, long long R, int D,
                                        float eps, float* __restrict__ mean, float* __restrict__ std) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  mean[c] = (float)(sum[c] / (double)R);
  std[c] = (float)sqrt(ss[c] / (double)(R - 1)) + eps;
}

template <typename TO>
__device__ __forceinline__ void store_out(TO* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// one row of D floats: out = (x - mean) / std (IEEE division, like torch), zero, or a plain copy; 16-byte accesses when
// the width allows
template <typename TO>
__device__ __forceinline__ void emit_row(TO* __restrict__ dst, const float* __restrict__ src, const float* __restrict__ mean,
                                         const float* __restrict__ sd, int D, int lane, bool vec) {
  if (vec) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src != nullptr) {
        v = *reinterpret_cast<const float4*>(src + c);
        if (mean != nullptr) {
          const float4 m = *reinterpret_cast<const float4*>(mean + c), s = *reinterpret_cast<const float4*>(sd + c);
          v.x = __fdiv_rn(v.x - m.x, s.x); v.y = __fdiv_rn(v.y - m.y, s.y);
          v.z = __fdiv_rn(v.z - m.z, s.z); v.w = __fdiv_rn(v.w - m.w, s.w);
        }
      }
      if (sizeof(TO) == 4) {
        *reinterpret_cast<float4*>(dst + c) = v;
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dst + c) = pk;
      }
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float v = 0.f;
      if (src != nullptr) v = mean != nullptr ? __fdiv_rn(src[c] - mean[c], sd[c]) : src[c];
      store_out(dst + c, v);
    }
  }
}

// One CTA row-group per padded (b, t) row: rows [0, B*Tmax) are video rows, rows [B*Tmax, B*Tmax + B) the audio rows.
template <typename TO>
__global__ void __launch_bounds__(256)
collate_kernel(const float* __restrict__ frames, const long long* __restrict__ offsets, const float* __restrict__ audio,
               const long long* __restrict__ labels, const long long* __restrict__ idx, const float* __restrict__ mean_v,
               const float* __restrict__ std_v, const float* __restrict__ mean_a, const float* __restrict__ std_a,
               TO* __restrict__ video_out, TO* __restrict__ audio_out, long long* __restrict__ labels_out,
               uint8_t* __restrict__ mask_out, int B, int Tmax, int Dv, int Da) {
  const long long rows_v = (long long)B * Tmax;
  const long long rows = rows_v + B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    if (row < rows_v) {
      const int b = (int)(row / Tmax), t = (int)(row % Tmax);
      const long long s = idx[b];
      const long long f0 = offsets[s], len = offsets[s + 1] - f0;
      const bool real = t < len;
      if (lane == 0) mask_out[row] = real ? 0 : 1;          // True = padded (pad_sequence(..., padding_value=True))
      // padded rows: zeros (pad_sequence(..., padding_value=0.0))
      emit_row(video_out + row * Dv, real ? frames + (f0 + t) * Dv : nullptr, mean_v, std_v, Dv, lane, (Dv & 3) == 0);
    } else {
      const int b = (int)(row - rows_v);
      const long long s = idx[b];
      emit_row(audio_out + (long long)b * Da, audio + s * Da, mean_a, std_a, Da, lane, (Da & 3) == 0);
      if (lane == 0 && labels_out != nullptr) labels_out[b] = labels[s];
    }
  }
}

// ---- bf16-resident variant: the features are z-scored and rounded to bf16 ONCE (normalize_rows_kernel), after which a
// batch is a pure gather of 2-byte rows -- half the bytes per batch and per resident sample; the values are the same
// bits the fp32-resident path produces with bf16 output (same arithmetic, one rounding)
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ sd,
                      bf16* __restrict__ out, long long R, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < R; row += (long long)gridDim.x * 8)
    emit_row(out + row * D, x + row * D, mean, sd, D, lane, (D & 3) == 0);
}

__device__ __forceinline__ void copy_row_bf16(bf16* __restrict__ dst, const bf16* __restrict__ src, int D, int lane, bool vec) {
  if (vec) {   // D % 8 == 0: 16-byte accesses
    for (int c = lane * 8; c < D; c += 256)
      *reinterpret_cast<uint4*>(dst + c) = src != nullptr ? *reinterpret_cast<const uint4*>(src + c) : make_uint4(0u, 0u, 0u, 0u);
  } else {
    for (int c = lane; c < D; c += 32) dst[c] = src != nullptr ? src[c] : __float2bfloat16_rn(0.f);
  }
}

__global__ void __launch_bounds__(256)
collate_bf16_kernel(const bf16* __restrict__ frames, const long long* __restrict__ offsets, const bf16* __restrict__ audio,
                    const long long* __restrict__ labels, const long long* __restrict__ idx, bf16* __restrict__ video_out,
                    bf16* __restrict__ audio_out, long long* __restrict__ labels_out, uint8_t* __restrict__ mask_out, int B,
                    int Tmax, int Dv, int Da) {
  const long long rows_v = (long long)B * Tmax;
  const long long rows = rows_v + B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    if (row < rows_v) {
      const int b = (int)(row / Tmax), t = (int)(row % Tmax);
      const long long s = idx[b];
      const long long f0 = offsets[s], len = offsets[s + 1] - f0;
      const bool real = t < len;
      if (lane == 0) mask_out[row] = real ? 0 : 1;
      copy_row_bf16(video_out + row * Dv, real ? frames + (f0 + t) * Dv : nullptr, Dv, lane, (Dv & 7) == 0);
    } else {
      const int b = (int)(row - rows_v);
      const long long s = idx[b];
      copy_row_bf16(audio_out + (long long)b * Da, audio + s * Da, Da, lane, (Da & 7) == 0);
      if (lane == 0 && labels_out != nullptr) labels_out[b] = labels[s];
    }
  }
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_feature_stats(const float* x, int64_t R, int64_t D, float eps, float* mean, float* std, double* scratch,
                       void* stream) {
  MMER_CHECK_ARG(x && mean && std && scratch, "feature_stats: null pointer");
  MMER_CHECK_ARG(R >= 1 && D >= 1 && D <= (1 << 20), "feature_stats: bad shape R=%lld D=%lld", (long long)R, (long long)D);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * (size_t)D, st);
  if (e != cudaSuccess) return cuda_fail(e, "memset(feature_stats)");
  const unsigned gx = (unsigned)((D + 255) / 256);
  long long gy = (long long)sm_count() * 8 / gx;
  if (gy < 1) gy = 1;
  if (gy > R) gy = R;
  dim3 grid(gx, (unsigned)gy);
  feature_moment_kernel<<<grid, 256, 0, st>>>(x, R, (int)D, nullptr, 0, scratch);
  MMER_LAUNCH_CHECK("feature_moment_kernel(sum)");
  feature_moment_kernel<<<grid, 256, 0, st>>>(x, R, (int)D, scratch, 1, scratch + D);
  MMER_LAUNCH_CHECK("feature_moment_kernel(ss)");
  feature_finalize_kernel<<<gx, 256, 0, st>>>(scratch, scratch + D, R, (int)D, eps, mean, std);
  MMER_LAUNCH_CHECK("feature_finalize_kernel");
  return 0;
}

int mmer_normalize_rows(const float* x, const float* mean, const float* std, void* out_bf16, int64_t R, int64_t D,
                        void* stream) {
  MMER_CHECK_ARG(x && out_bf16 && (mean == nullptr) == (std == nullptr), "normalize_rows: bad pointers");
  MMER_CHECK_ARG(R >= 0 && D >= 1, "normalize_rows: bad shape");
  if (R == 0) return 0;
  long long grid = (R + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  normalize_rows_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, mean, std, (bf16*)out_bf16, R, (int)D);
  MMER_LAUNCH_CHECK("normalize_rows_kernel");
  return 0;
}

int mmer_collate_bf16(const void* frames, const int64_t* offsets, const void* audio, const int64_t* labels,
                      const int64_t* idx, void* video_out, void* audio_out, int64_t* labels_out, uint8_t* mask_out, int64_t B,
                      int64_t Tmax, int64_t Dv, int64_t Da, void* stream) {
  MMER_CHECK_ARG(frames && offsets && audio && idx && video_out && audio_out && mask_out, "collate_bf16: null pointer");
  MMER_CHECK_ARG(labels_out == nullptr || labels != nullptr, "collate_bf16: labels_out needs labels");
  MMER_CHECK_ARG(B >= 0 && Tmax >= 0 && Dv >= 1 && Da >= 1, "collate_bf16: bad shape");
  if (B == 0) return 0;
  const long long rows = B * Tmax + B;
  long long grid = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  typedef const long long* LP;
  collate_bf16_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)frames, (LP)offsets, (const bf16*)audio,
                                                                       (LP)labels, (LP)idx, (bf16*)video_out, (bf16*)audio_out,
                                                                       (long long*)labels_out, mask_out, (int)B, (int)Tmax,
                                                                       (int)Dv, (int)Da);
  MMER_LAUNCH_CHECK("collate_bf16_kernel");
  return 0;
}

int mmer_collate(const float* frames, const int64_t* offsets, const float* audio, const int64_t* labels, const int64_t* idx,
                 const float* mean_v, const float* std_v, const float* mean_a, const float* std_a, void* video_out,
                 void* audio_out, int64_t* labels_out, uint8_t* mask_out, int64_t B, int64_t Tmax, int64_t Dv, int64_t Da,
                 int out_dtype, void* stream) {
  MMER_CHECK_ARG(frames && offsets && audio && idx && video_out && audio_out && mask_out, "collate: null pointer");
  MMER_CHECK_ARG((mean_v == nullptr) == (std_v == nullptr) && (mean_a == nullptr) == (std_a == nullptr),
                 "collate: mean and std come together");
  MMER_CHECK_ARG(labels_out == nullptr || labels != nullptr, "collate: labels_out needs labels");
  MMER_CHECK_ARG(B >= 0 && Tmax >= 0 && Dv >= 1 && Da >= 1, "collate: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = B * Tmax + B;
  long long grid = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  typedef const long long* LP;
  if (out_dtype == MMER_F32)
    collate_kernel<float><<<(unsigned)grid, 256, 0, st>>>(frames, (LP)offsets, audio, (LP)labels, (LP)idx, mean_v, std_v, mean_a,
                                                          std_a, (float*)video_out, (float*)audio_out, (long long*)labels_out,
                                                          mask_out, (int)B, (int)Tmax, (int)Dv, (int)Da);
  else if (out_dtype == MMER_BF16)
    collate_kernel<bf16><<<(unsigned)grid, 256, 0, st>>>(frames, (LP)offsets, audio, (LP)labels, (LP)idx, mean_v, std_v, mean_a,
                                                         std_a, (bf16*)video_out, (bf16*)audio_out, (long long*)labels_out,
                                                         mask_out, (int)B, (int)Tmax, (int)Dv, (int)Da);
  else
    MMER_CHECK_ARG(false, "collate: unsupported output dtype");
  MMER_LAUNCH_CHECK("collate_kernel");
  return 0;
}

}  // extern "C"
