// Short-sequence attention forward, float activations: explicit instantiations (split for build time).
#include "attention_small.cuh"

namespace mmer {

int mha_fwd_small_f32(int d, int SP, const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H,
                       DropCfg dc, cudaStream_t st) {
  return d == 64 ? mha_fwd_sp<float, 64>(SP, qkv, mask, out, probs, B, Tn, H, dc, st)
                 : mha_fwd_sp<float, 32>(SP, qkv, mask, out, probs, B, Tn, H, dc, st);
}

}  // namespace mmer
