// Short-sequence attention backward, bf16 activations: explicit instantiations (split for build time).
#include "attention_small.cuh"

namespace mmer {

int mha_bwd_small_bf16(int d, int SP, const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn,
                       int H, DropCfg dc, cudaStream_t st) {
  return d == 64 ? mha_bwd_sp<bf16, 64>(SP, qkv, mask, dout, dqkv, B, Tn, H, dc, st)
                 : mha_bwd_sp<bf16, 32>(SP, qkv, mask, dout, dqkv, B, Tn, H, dc, st);
}

}  // namespace mmer
