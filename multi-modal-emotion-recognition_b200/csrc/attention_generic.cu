// Long-sequence attention (S = T+1 > 32, e.g. the T=256 interpretability configuration).
#include "common.cuh"

namespace mmer {

int mha_fwd_generic(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H,
                    int64_t d, int dtype, DropCfg dc, cudaStream_t st) {
  (void)qkv; (void)mask; (void)out; (void)probs; (void)B; (void)H; (void)d; (void)dtype; (void)dc; (void)st;
  set_error("mha_fwd: sequences longer than 32 tokens (T=%lld) are not implemented yet", (long long)T);
  return MMER_ERR_UNSUPPORTED;
}
int mha_bwd_generic(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int64_t B, int64_t T,
                    int64_t H, int64_t d, int dtype, DropCfg dc, cudaStream_t st) {
  (void)qkv; (void)mask; (void)dout; (void)dqkv; (void)B; (void)H; (void)d; (void)dtype; (void)dc; (void)st;
  set_error("mha_bwd: sequences longer than 32 tokens (T=%lld) are not implemented yet", (long long)T);
  return MMER_ERR_UNSUPPORTED;
}

}  // namespace mmer
