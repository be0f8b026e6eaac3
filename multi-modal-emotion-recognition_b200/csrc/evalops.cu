// On-device evaluation bookkeeping: what the reference's validation / test loops do per batch with
// .item() / .cpu() round trips (train2.py:593-607, 651-667, 724-741) -- `_, predicted = torch.max(probs, dim=1)`,
// `correct += (predicted == labels).sum()`, extend(all_preds / all_labels) -- as one kernel that accumulates a confusion
// matrix in device memory.  Accuracy, macro / micro precision / recall / F1 and sklearn's confusion_matrix all follow
// from that matrix on the host at the end of the epoch (one 288-byte copy instead of two syncs per batch).
#include "common.cuh"

namespace mmer {

constexpr int EVAL_MAXC = 16;

__global__ void __launch_bounds__(256)
eval_accumulate_kernel(const float* __restrict__ probs, const long long* __restrict__ labels,
                       long long* __restrict__ predicted, unsigned long long* __restrict__ conf, int B, int C) {
  __shared__ unsigned int sconf[EVAL_MAXC * EVAL_MAXC];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sconf[i] = 0u;
  __syncthreads();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const float* p = probs + (long long)b * C;
    int best = 0;
    float bv = p[0];
    for (int c = 1; c < C; ++c) {
      const float v = p[c];
      if (v > bv || (v != v && bv == bv)) { bv = v; best = c; }   // first maximum; NaN wins like torch.max
    }
    if (predicted != nullptr) predicted[b] = best;
    const long long y = labels[b];
    if (y >= 0 && y < C) atomicAdd(&sconf[(int)y * C + best], 1u);   // rows: true class, columns: predicted class
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (sconf[i]) atomicAdd(conf + i, (unsigned long long)sconf[i]);
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int mmer_eval_accumulate(const float* probs, const int64_t* labels, int64_t* predicted, int64_t* confusion, int64_t B,
                         int64_t C, void* stream) {
  MMER_CHECK_ARG(probs && labels && confusion, "eval_accumulate: null pointer");
  MMER_CHECK_ARG(C >= 1 && C <= EVAL_MAXC, "eval_accumulate: at most %d classes", EVAL_MAXC);
  if (B <= 0) return 0;
  long long grid = (B + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (grid > cap) grid = cap;
  eval_accumulate_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(probs, (const long long*)labels, (long long*)predicted,
                                                                          (unsigned long long*)confusion, (int)B, (int)C);
  MMER_LAUNCH_CHECK("eval_accumulate_kernel");
  return 0;
}

}  // extern "C"
