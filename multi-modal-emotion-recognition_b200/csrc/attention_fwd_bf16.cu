// Short-sequence attention forward, bf16 activations: explicit instantiations (split for build time).
#include "attention_small.cuh"

namespace mmer {

int mha_fwd_small_bf16(int d, int SP, const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                       DropCfg dc, cudaStream_t st) {
  return d == 64 ? mha_fwd_sp<bf16, 64>(SP, qkv, mask, out, probs, B, Tn, H, dc, st)
                 : mha_fwd_sp<bf16, 32>(SP, qkv, mask, out, probs, B, Tn, H, dc, st);
}

}  // namespace mmer
