// Multi-head self-attention entry points: dispatch on sequence length, head size and dtype.
// Short sequences (S = T+1 <= 32): bf16 -> attention_mma.cu (one CTA per sample, bulk-copied rows, warp-level
// tensor-core MMAs); fp32 parity mode -> attention_small.cuh (one warp per (sample, head), register tiled FMA).
// Longer sequences: attention_generic.cu (one CTA per (sample, head), K/V resident in shared memory).
#include "common.cuh"

namespace mmer {

int mha_fwd_small_bf16(int d, int SP, const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                       DropCfg dc, cudaStream_t st);
int mha_fwd_small_f32(int d, int SP, const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                      DropCfg dc, cudaStream_t st);
int mha_bwd_small_bf16(int d, int SP, const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn,
                       int H, DropCfg dc, cudaStream_t st);
int mha_bwd_small_f32(int d, int SP, const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn,
                      int H, DropCfg dc, cudaStream_t st);
int mha_fwd_mma(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, int d, DropCfg dc,
                cudaStream_t st);
int mha_bwd_mma(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B, int Tn, int H,
                int d, DropCfg dc, cudaStream_t st);
extern int g_debug[16];
bool mha_long_supported(int Tn, int d, int dtype);
int mha_fwd_long(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, DropCfg dc, cudaStream_t st,
                 float* lse);
int mha_bwd_long(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H, DropCfg dc,
                 cudaStream_t st, const float* lse, const void* fwd_out);
int mha_fwd_generic(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H,
                    int64_t d, int dtype, DropCfg dc, cudaStream_t st);
int mha_bwd_generic(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int64_t B, int64_t T,
                    int64_t H, int64_t d, int dtype, DropCfg dc, cudaStream_t st);

// Engine-side entry points: the same dispatch plus, for the long-sequence tensor-core kernels, the softmax statistics
// (max, 1 / sum per query row; [B, H, S, 2] fp32) that forward can hand to backward together with its own output.
int mha_fwd_ex(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H, int64_t d,
               int dtype, float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, float* lse) {
  if (lse != nullptr && B > 0 && T + 1 > 32 && mha_long_supported((int)T, (int)d, dtype) && !g_debug[MMER_DEBUG_ATT_SIMT])
    return mha_fwd_long(qkv, mask, out, probs, (int)B, (int)T, (int)H, make_drop(drop_p, seed, site), st, lse);
  return mmer_mha_fwd(qkv, mask, out, probs, B, T, H, d, dtype, drop_p, seed, site, (void*)st);
}
int mha_bwd_ex(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias_qkv, int64_t B, int64_t T,
               int64_t H, int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, const float* lse,
               const void* fwd_out) {
  if (lse != nullptr && fwd_out != nullptr && B > 0 && T + 1 > 32 && mha_long_supported((int)T, (int)d, dtype) &&
      !g_debug[MMER_DEBUG_ATT_SIMT]) {
    MMER_TRY(mha_bwd_long(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, make_drop(drop_p, seed, site), st, lse, fwd_out));
    if (dbias_qkv != nullptr) return mmer_colsum(dqkv, dbias_qkv, B * (T + 1), 3 * H * d, 3 * H * d, dtype, (void*)st);
    return 0;
  }
  return mmer_mha_bwd(qkv, mask, dout, dqkv, dbias_qkv, B, T, H, d, dtype, drop_p, seed, site, (void*)st);
}

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
  if (T + 1 > 32 && mha_long_supported((int)T, (int)d, dtype) && !g_debug[MMER_DEBUG_ATT_SIMT])
    return mha_fwd_long(qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st, nullptr);   // tensor-core tiles, S <= 384
  if (T + 1 > 32) return mha_fwd_generic(qkv, mask, out, probs, B, T, H, d, dtype, dc, st);
  const int SP = (int)((T + 1 + 3) & ~3LL);
  if (dtype == MMER_BF16 && !g_debug[MMER_DEBUG_ATT_SIMT])
    return mha_fwd_mma(qkv, mask, out, probs, (int)B, (int)T, (int)H, (int)d, dc, st);
  if (dtype == MMER_BF16) return mha_fwd_small_bf16((int)d, SP, qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
  return mha_fwd_small_f32((int)d, SP, qkv, mask, out, probs, (int)B, (int)T, (int)H, dc, st);
}

int mmer_mha_bwd(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias_qkv, int64_t B,
                 int64_t T, int64_t H, int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  MMER_CHECK_ARG(qkv && dout && dqkv, "mha_bwd: null pointer");
  MMER_CHECK_ARG(d == 64 || d == 32, "mha_bwd: head dim %lld unsupported (32 or 64)", (long long)d);
  MMER_CHECK_ARG(T >= 1 && H >= 1, "mha_bwd: bad shape");
  if (B <= 0) return 0;
  DropCfg dc = make_drop(drop_p, seed, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (T + 1 <= 32 && dtype == MMER_BF16 && !g_debug[MMER_DEBUG_ATT_SIMT])   // in_proj bias gradient fused
    return mha_bwd_mma(qkv, mask, dout, dqkv, dbias_qkv, (int)B, (int)T, (int)H, (int)d, dc, st);
  if (T + 1 > 32 && mha_long_supported((int)T, (int)d, dtype) && !g_debug[MMER_DEBUG_ATT_SIMT]) {
    MMER_TRY(mha_bwd_long(qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st, nullptr, nullptr));
  } else if (T + 1 > 32) {
    MMER_TRY(mha_bwd_generic(qkv, mask, dout, dqkv, B, T, H, d, dtype, dc, st));
  } else {
    const int SP = (int)((T + 1 + 3) & ~3LL);
    if (dtype == MMER_BF16) MMER_TRY(mha_bwd_small_bf16((int)d, SP, qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st));
    else MMER_TRY(mha_bwd_small_f32((int)d, SP, qkv, mask, dout, dqkv, (int)B, (int)T, (int)H, dc, st));
  }
  if (dbias_qkv != nullptr) return mmer_colsum(dqkv, dbias_qkv, B * (T + 1), 3 * H * d, 3 * H * d, dtype, stream);
  return 0;
}

}  // extern "C"
