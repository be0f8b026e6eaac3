/* mmer.h -- C ABI of libmmer_sm100.so: the B200 (sm_100a) kernels behind the audio-visual
 * fusion classifier step of EvanZJ/multi-modal-emotion-recognition.
 *
 * The reference has no native layer; its hot path is PyTorch module calls.  Each entry
 * point below names the reference call site (file:line, relative to the reference repo)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 *
 * Conventions:
 *   - plain pointers and sizes; all pointers are DEVICE pointers unless stated otherwise
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocates, never synchronises, owns nothing
 *   - returns 0 on success, a negative MMER_ERR_* code otherwise; the message is
 *     available from mmer_last_error() (thread local)
 *   - the device is the caller's current device
 *   - `dtype` selects the storage type of activations: MMER_F32 or MMER_BF16.  All
 *     statistics, reductions, losses, gradients of parameters and optimizer state are fp32.
 *   - token matrices are batch-major: row = b * S + s, S = T + 1 (audio token last).
 */
#ifndef MMER_H_
#define MMER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMER_VERSION 100

#define MMER_OK 0
#define MMER_ERR_ARG (-1)      /* bad argument / unsupported shape */
#define MMER_ERR_CUDA (-2)     /* CUDA runtime or driver error */
#define MMER_ERR_UNSUPPORTED (-3)

#define MMER_F32 0
#define MMER_BF16 1

#define MMER_MAJOR_K 0   /* reduction dim contiguous: matrix stored [rows][K]  */
#define MMER_MAJOR_MN 1  /* row dim contiguous:      matrix stored [K][rows]  */

#define MMER_LOSS_FOCAL 0   /* FocalLoss, train.py:20-37 == train2.py:40-70          */
#define MMER_LOSS_WCE 1     /* nn.CrossEntropyLoss(weight=w), train2.py:523          */
#define MMER_REDUCE_MEAN 0
#define MMER_REDUCE_SUM 1
#define MMER_REDUCE_NONE 2

int mmer_version(void);
const char* mmer_last_error(void);
/* debug knobs used by the GPU tests (descriptor variants, forcing the SIMT GEMM ...) */
#define MMER_DEBUG_MN_SWAP 0     /* swap LBO/SBO in MN-major UMMA descriptors */
#define MMER_DEBUG_FORCE_BN 1    /* force the tcgen05 GEMM N tile (128 or 256) */
#define MMER_DEBUG_FORCE_SIMT 2  /* route bf16 GEMMs through the fp32 FMA kernel (debug only) */
#define MMER_DEBUG_DIRECT_STORE 3 /* tcgen05 GEMM: bypass the smem + TMA-store epilogue */
#define MMER_DEBUG_NO_PAIR 5     /* tcgen05 GEMM: never use CTA pairs (cta_group::2); A/B timing */
#define MMER_DEBUG_GENERIC_EPI 6 /* tcgen05 GEMM: always the run-time-flag epilogue, never a specialised one; A/B timing */
#define MMER_DEBUG_ATT_ROWS 7    /* bf16 short-sequence attention, d=64: CTA-per-sample bulk-row kernels instead of the warp-pipelined TMA-tile ones */
#define MMER_DEBUG_NO_PDL 8      /* launch every kernel fully serialised (no programmatic dependent launch); A/B timing */
#define MMER_DEBUG_RESERVE_SMS 9 /* size every persistent grid for (SMs - value): room for an NCCL kernel beside the step */
#define MMER_DEBUG_FORCE_SPLITS 10 /* tcgen05 GEMM, accumulate mode: force the split-K factor; A/B timing */
#define MMER_DEBUG_LN_VARIANT 12 /* fused GEMM+LayerNorm kernel: switch parts of the epilogue off (bit mask); A/B timing only, results are wrong */
#define MMER_DEBUG_NO_LN_FUSE 11 /* engine, post-norm sub-layer tails: 0 default (GEMM with residual epilogue -> z, then LayerNorm(z)); 1 GEMM -> a, add_ln(x, a); 2 the fully fused GEMM+LayerNorm kernel; A/B timing */
#define MMER_DEBUG_SERVE_GLOBAL 13 /* batch-1 serving forward, S <= 8: 0 head-local / split-K kernel (serve_small.cu); 1 output-feature split with the L2 exchange (serve.cu; run-to-run bit-reproducible); 2 shared-memory broadcast (serve_dsmem.cu); A/B timing */
#define MMER_DEBUG_SERVE_STAMPS 14 /* serve_small.cu: extra globaltimer marks inside phase 0 and layer 0 (they cost time) */
#define MMER_DEBUG_EMBED_GENERIC 15 /* token assembly (LayerNorm variant): always the general tile kernels + embed_dpos, never the position-stable ones; A/B timing */
#define MMER_DEBUG_ATT_SIMT 4    /* bf16 short-sequence attention: use the FMA kernels instead of the MMA ones (A/B timing) */
int mmer_debug_set(int key, int value);
int mmer_debug_get(int key);
/* number of kernels this library has launched in the current process (bench accounting) */
int64_t mmer_launch_count(void);

/* ------------------------------------------------------------------------------------
 * GEMM: D[M,N] = epilogue( A[M,K] . B[N,K]^T ).  Replaces every nn.Linear forward on the
 * path (train2.py:150,153,217-228; the in_proj / out_proj / linear1 / linear2 of
 * nn.TransformerEncoderLayer configured at train2.py:111-118, train.py:54-57) and their
 * autograd dgrad / wgrad.  in_dtype MMER_BF16 runs on tcgen05 tensor cores (TMA-fed,
 * accumulators in TMEM); in_dtype MMER_F32 runs an fp32 FMA kernel (parity mode).
 *   forward : A = X [M,K] K-major,  B = W [N,K] K-major
 *   dgrad   : A = dY [M,N'] K-major, B = W viewed as [K',N'] -> MMER_MAJOR_MN
 *   wgrad   : A = dY viewed [N',M] MN-major, B = X viewed [K',M] MN-major, accumulate=1
 * epilogue order: +bias, relu, *dropout, *gate, +residual.
 * ---------------------------------------------------------------------------------- */
typedef struct mmer_gemm_args {
  const void* A;
  const void* B;
  void* D;
  const float* bias;    /* [N] fp32 or NULL */
  const void* residual; /* [M,N] out_dtype, leading dim ldd, or NULL */
  const void* gate;     /* [M,N] out_dtype, leading dim ldd, or NULL: acc *= gate>0 ? gate_scale : 0 */
  int64_t M, N, K;
  int64_t lda, ldb, ldd; /* row stride, in elements, of each matrix as stored */
  int32_t a_major, b_major;
  int32_t in_dtype, out_dtype;
  int32_t accumulate; /* D += result; requires out_dtype MMER_F32 (split-K, fp32 atomics) */
  int32_t relu;
  float drop_p;
  float gate_scale;
  uint64_t seed;
  uint32_t drop_site;
  uint32_t reserved;
  float* a_rowsum; /* optional fp32 [M]: += sum_k A[m][k].  Needs accumulate != 0 and an MN-major A: in a weight-gradient
                    * GEMM (A = dY^T) this is the bias gradient of the same Linear, produced by the same kernel */
  /* 1 bit per output element (row-major, N/8 bytes per row; bf16 output, N % 64 == 0, ldd == N):
   * relu_mask_out - written by a forward call: bit = stored value > 0 (after ReLU and dropout);
   * gate_bits     - read by the matching dgrad call: acc *= bit ? gate_scale : 0.  16x less traffic than `gate`. */
  uint8_t* relu_mask_out;
  const uint8_t* gate_bits;
  /* optional fp32 [N]: += column sums of D as STORED (bf16 output through the staged epilogue; any other output path
   * adds them with a separate pass).  For a data-gradient GEMM dX = dY W this is the bias gradient of the Linear that
   * produced X -- linear1's bias gradient comes out of linear2's dgrad epilogue, where the gradient tile is in shared
   * memory anyway, instead of a row-sum MMA in linear1's weight-gradient GEMM. */
  float* d_colsum;
} mmer_gemm_args;
int mmer_gemm(const mmer_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * Token assembly: LayerNorm of both projections, concat, +pos_embed, dropout.
 * train2.py:151,154,157-161.  pv [B*T,F], pa [B,F] are the projection outputs;
 * x0 [B*S,F]; stats [B*S,2] fp32 (mean, rstd).  pos is fp32 [>=S,F].
 * ---------------------------------------------------------------------------------- */
int mmer_embed_fwd(const void* pv, const void* pa, const float* gv, const float* bv, const float* ga,
                   const float* ba, const float* pos, void* x0, float* stats, int64_t B, int64_t T,
                   int64_t F, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);
/* backward: dpv, dpa get the LayerNorm input gradients; dgv.. dpos are ACCUMULATED (+=) */
/* dbias_v / dbias_a (optional, fp32 [F]): += column sums of dpv / dpa as stored, i.e. the bias gradients of the two
 * input projections (LayerNorm variant only; NULL otherwise). */
int mmer_embed_bwd(const void* dx0, const void* pv, const void* pa, const float* stats, const float* gv,
                   const float* ga, void* dpv, void* dpa, float* dgv, float* dbv, float* dga, float* dba,
                   float* dpos, float* dbias_v, float* dbias_a, int64_t B, int64_t T, int64_t F, int dtype,
                   float drop_p, uint64_t seed, uint32_t site, void* stream);

/* ------------------------------------------------------------------------------------
 * Row kernel: z = x + dropout_a(a);  y = dropout_y(relu?(LayerNorm(z))).
 * Encoder: x + dropout(sublayer) then norm1/norm2 (post-norm nn.TransformerEncoderLayer);
 * head: Linear -> LayerNorm -> ReLU -> Dropout (train2.py:217-226).  x may be NULL (z = a).
 * stats [M,2] fp32.  gamma NULL => no normalisation (identity).
 * ---------------------------------------------------------------------------------- */
int mmer_add_ln_fwd(const void* x, const void* a, const float* gamma, const float* beta, void* y,
                    float* stats, int64_t M, int64_t F, int dtype, int relu, float drop_a_p,
                    uint32_t site_a, float drop_y_p, uint32_t site_y, uint64_t seed, void* stream);
/* backward.  dz = gradient w.r.t. z (goes to the residual stream x); da = dz * mask_a,
 * written only when drop_a_p > 0 and da != NULL (otherwise da == dz).  dgamma/dbeta and
 * dbias (column sum of da, i.e. the bias gradient of the Linear that produced `a`) are
 * ACCUMULATED; each may be NULL.  Nothing of the forward output is re-read: the ReLU sign
 * and both dropout masks are recomputed from x, a, stats, gamma, beta and the seed. */
int mmer_add_ln_bwd(const void* dy, const void* x, const void* a, const float* stats, const float* gamma,
                    const float* beta, void* dz, void* da, float* dgamma, float* dbeta, float* dbias,
                    int64_t M, int64_t F, int dtype, int relu, float drop_a_p, uint32_t site_a,
                    float drop_y_p, uint32_t site_y, uint64_t seed, void* stream);
/* The same backward after the FUSED forward below: `z` is the stored pre-LayerNorm sum x + dropout(sub-layer output)
 * (what mmer_gemm_ln_fwd writes), so x and the sub-layer output are not re-read; the dropout mask (drop_a_p, site_a,
 * seed) is regenerated only for da = dz o mask.  da may be NULL when drop_a_p == 0 (then da == dz). */
int mmer_add_ln_bwd_z(const void* dy, const void* z, const float* stats, const float* gamma, void* dz, void* da,
                      float* dgamma, float* dbeta, float* dbias, int64_t M, int64_t F, int dtype, float drop_a_p,
                      uint32_t site_a, uint64_t seed, void* stream);
/* Post-norm encoder sub-layer tail in ONE tcgen05 kernel (train2.py:111-118 -> nn.TransformerEncoderLayer,
 * norm_first=False: `x = norm1(x + dropout1(sa_block(x)))`, `x = norm2(x + dropout2(ff_block(x)))`):
 *   z = residual + dropout(a[M,K] . w[512,K]^T + bias)      y = LayerNorm(z) * gamma + beta      (bf16, N = 512)
 * z_out [M,512] bf16 (input of mmer_add_ln_bwd_z), y_out [M,512] bf16, stats [M,2] fp32 (mean, rstd).  residual may be
 * NULL.  y is computed from the bf16-rounded z, i.e. from exactly what backward reads.  Dropout mask = the counter hash
 * of (seed, site, row * 512 + col), the same one mmer_add_ln_fwd / _bwd use. */
int mmer_gemm_ln_fwd(const void* a, const void* w, const float* bias, const void* residual, const float* gamma,
                     const float* beta, void* z_out, void* y_out, float* stats, int64_t M, int64_t K, float drop_p,
                     uint32_t site, uint64_t seed, void* stream);

/* Masked mean pooling over the S tokens of each sample + out_norm (train2.py:184-191;
 * train.py:100-104 without the norm: gamma == NULL).  mask [B,T] bytes, 1 = padded, or NULL. */
int mmer_pool_ln_fwd(const void* x, const uint8_t* mask, const float* gamma, const float* beta,
                     float* pooled, void* fused, float* stats, int64_t B, int64_t T, int64_t F, int dtype,
                     void* stream);
int mmer_pool_ln_bwd(const void* dfused, const float* pooled, const float* stats, const float* gamma,
                     const uint8_t* mask, void* dx, float* dgamma, float* dbeta, int64_t B, int64_t T,
                     int64_t F, int dtype, void* stream);

/* out[n] += sum_m x[m,n]  (bias gradients) */
int mmer_colsum(const void* x, float* out, int64_t M, int64_t N, int64_t ldx, int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Multi-head self-attention over [T video tokens ; 1 audio token] with key padding mask.
 * Replaces nn.MultiheadAttention inside the encoder (called via train2.py:173-176).
 * qkv [B*S, 3*H*d] packed as torch's in_proj output (q | k | v, heads contiguous);
 * out [B*S, H*d].  probs (optional, fp32 [B,H,S,S]) receives the softmax weights: this is
 * the defined attention-weight output for `return_attn=True`.
 * ---------------------------------------------------------------------------------- */
int mmer_mha_fwd(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T,
                 int64_t H, int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);
/* dbias_qkv (optional, fp32 [3*H*d]): += column sums of dqkv = gradient of in_proj_bias (fused into the kernel). */
int mmer_mha_bwd(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias_qkv, int64_t B,
                 int64_t T, int64_t H, int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);

/* Final Linear(hidden -> C) + softmax (train2.py:228,290; train.py:128-129), C <= 16.
 * logits/probs fp32 [B,C]. */
int mmer_head_out_fwd(const void* h, const float* W, const float* b, float* logits, float* probs,
                      int64_t B, int64_t K, int64_t C, int dtype, void* stream);
/* dh [B,K] (dtype); dW [C,K], db [C] accumulated. */
int mmer_head_out_bwd(const float* dlogits, const void* h, const float* W, void* dh, float* dW, float* db,
                      int64_t B, int64_t K, int64_t C, int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Loss forward + gradient in one kernel.  FocalLoss (train.py:27-37): alpha optional,
 * plain mean over B.  Weighted CE (train2.py:523,572): sum(w_y ce)/sum(w_y).
 * loss_out: 1 fp32 (zeroed by the call; for REDUCE_NONE unused), per_sample [B] optional,
 * dlogits [B,C] optional (= grad_scale * dloss/dlogits).  scratch: >= 2 fp32.
 * ---------------------------------------------------------------------------------- */
int mmer_loss_fwd_bwd(const float* logits, const int64_t* labels, const float* alpha, int kind, float gamma,
                      int reduction, float* loss_out, float* per_sample, float* dlogits, float* scratch,
                      int64_t B, int64_t C, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * optim.Adam(lr, weight_decay) step over a flat fp32 buffer (train.py:252,297;
 * train2.py:525,578) with optional fused clip_grad_norm_ (train2.py:576):
 *   g' = g * grad_scale * min(1, max_norm / (sqrt(*sumsq * grad_scale^2) + 1e-6))
 *   g' += wd * p;  m,v update;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
 * lr and step are runtime scalars (ReduceLROnPlateau, train2.py:526,614).  sumsq may be
 * NULL (no clipping).  shadow (bf16 copy of p) may be NULL.
 * ---------------------------------------------------------------------------------- */
int mmer_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                   const float* sumsq, float max_norm, void* stream);
/* Data-parallel variant over NVSwitch multicast (NVLS): gradient reduce-scatter + Adam on elements [lo, hi) + all-gather
 * of the updated weights in ONE kernel.  g_mc / p_mc / shadow_mc are MULTICAST addresses of symmetric buffers (every rank
 * maps the same layout): the kernel reads sum-over-ranks gradients with multimem.ld_reduce and writes the new fp32
 * weights (+ bf16 shadow) to all ranks with multimem.st; p_local is this rank's own full weight buffer, m / v hold the
 * moments of the rank's SHARD only (hi - lo elements, element lo first: ZeRO-1).  The reference has
 * no distributed code (train2.py:570-579 is single-process); this replaces NCCL all-reduce + mmer_adam_step for a
 * torchrun data-parallel job.  The caller issues a cross-rank barrier before (all gradients complete) and after (all
 * weights landed).  grad_scale = 1 / world_size.  lo, hi multiples of 4.
 * Gradient clipping (train2.py:576) in this mode: mmer_grad_sumsq_multicast sums the squares of the REDUCED gradient over
 * [lo, hi) into local_acc and publishes it into slot `rank` of a symmetric float array on every rank (slots_mc = its
 * multicast address); after a barrier, mmer_adam_step_multicast with sumsq_slots (this rank's copy of the array) and
 * n_slots = world size clips by min(1, max_norm / (sqrt(sum of slots) * grad_scale + 1e-6)), identical on all ranks.
 * sumsq_slots NULL = no clipping. */
int mmer_grad_sumsq_multicast(const float* g_mc, int64_t lo, int64_t hi, float* local_acc, float* slots_mc, int rank,
                              void* stream);
int mmer_adam_step_multicast(const float* p_local, float* p_mc, const float* g_mc, float* m, float* v, void* shadow_mc,
                             int64_t lo, int64_t hi, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int64_t step, float grad_scale, const float* sumsq_slots, int n_slots, float max_norm,
                             void* stream);
int mmer_grad_sumsq(const float* g, int64_t n, float* out /* zeroed by the call */, void* stream);
int mmer_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
int mmer_cast_f32(const void* src_bf16, float* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * BatchNorm1d of the train.py variant (train.py:51-52,66-74,116,125) over rows of [N,C].
 * training: batch statistics (biased variance), running stats updated with momentum 0.1
 * and unbiased variance; y = dropout(relu?(bn(x))).  stats_out is saved for backward.
 * ---------------------------------------------------------------------------------- */
int mmer_bn_fwd(const void* x, const float* gamma, const float* beta, float* running_mean,
                float* running_var, void* y, float* stats_out /* [2,C]: mean, rstd */, int64_t N, int64_t C,
                int dtype, int training, int relu, float momentum, float drop_p, uint64_t seed, uint32_t site,
                void* stream);
/* dgamma/dbeta ACCUMULATED; scratch >= 2*C fp32.  ReLU sign and dropout mask are recomputed. */
int mmer_bn_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                void* dx, float* dgamma, float* dbeta, float* scratch, int64_t N, int64_t C, int dtype,
                int training, int relu, float drop_p, uint64_t seed, uint32_t site, void* stream);

/* ------------------------------------------------------------------------------------
 * Whole-model entry points: MultimodalEmotionModel forward / backward / training step
 * (train2.py:281-292,570-579; train.py:139-142,293-297), one call each, all kernels
 * enqueued on `stream`.  See mmer_model in the engine section of DESIGN.md.
 * ---------------------------------------------------------------------------------- */
#define MMER_MAX_LAYERS 16
enum {
  MMER_G_POS = 0, MMER_G_WV, MMER_G_BV, MMER_G_WA, MMER_G_BA,
  MMER_G_NV_W, MMER_G_NV_B, MMER_G_NA_W, MMER_G_NA_B,   /* norm_video/norm_audio or bn_video/bn_audio */
  MMER_G_ON_W, MMER_G_ON_B,                             /* out_norm (v2 only) */
  MMER_G_C0_W, MMER_G_C0_B, MMER_G_C1_W, MMER_G_C1_B,   /* net.0 + net.1   | fc1 + bn_fc1 */
  MMER_G_C4_W, MMER_G_C4_B, MMER_G_C5_W, MMER_G_C5_B,   /* net.4 + net.5   | unused       */
  MMER_G_C8_W, MMER_G_C8_B,                             /* net.8           | fc2          */
  MMER_G_COUNT
};
enum {
  MMER_L_IN_W = 0, MMER_L_IN_B, MMER_L_OUT_W, MMER_L_OUT_B, MMER_L_FF1_W, MMER_L_FF1_B,
  MMER_L_FF2_W, MMER_L_FF2_B, MMER_L_N1_W, MMER_L_N1_B, MMER_L_N2_W, MMER_L_N2_B, MMER_L_COUNT
};

typedef struct mmer_model {
  int32_t variant;   /* 2: LayerNorm model (train2.py / back-end), 1: BatchNorm model (train.py) */
  int32_t dtype;     /* activation storage + GEMM input type */
  int32_t B, T;      /* batch, padded video length; S = T + 1 */
  int32_t video_dim, audio_dim, fused, heads, layers, ffn, hidden, classes;
  int32_t training;  /* dropout on, BatchNorm batch statistics */
  int32_t has_mask;
  float p_fusion, p_classifier;
  uint64_t seed;
  int64_t n_params;                       /* elements of the flat parameter buffer */
  int64_t off_g[MMER_G_COUNT];            /* element offsets into the flat buffers, -1 = absent */
  int64_t off_l[MMER_MAX_LAYERS][MMER_L_COUNT];
  /* buffers */
  float* params;        /* fp32 master weights, flat */
  void* shadow;         /* bf16 copy of params (dtype == MMER_BF16), flat, same offsets */
  float* grads;         /* fp32 flat, same offsets; backward ACCUMULATES into it */
  float* bn_state;      /* v1 only: running_mean/var for bn_video, bn_audio, bn_fc1: [2*F, 2*F, 2*H2] */
  void* workspace;      /* >= mmer_workspace_bytes() */
  int64_t workspace_bytes;
  /* per-call tensors */
  const void* video;    /* [B,T,video_dim] dtype */
  const void* audio;    /* [B,audio_dim]   dtype */
  const uint8_t* mask;  /* [B,T] 1 = padded, or NULL */
  float* logits;        /* [B,classes] fp32 out */
  float* probs;         /* [B,classes] fp32 out */
  void* fused_out;      /* [B,fused] dtype out, optional */
  float* attn_probs;    /* optional [layers,B,H,S,S] fp32 out (return_attn) */
  /* backward */
  const float* dlogits; /* [B,classes] fp32 */
  void* dvideo;         /* optional [B,T,video_dim] dtype: input gradients (Captum IG path) */
  void* daudio;         /* optional [B,audio_dim] dtype */
  /* partial evaluation, for callers that use the sub-modules on their own */
  int32_t stage;        /* 0 whole model; 1 CrossModalFusion only (train2.py:128-193); 2 EmotionClassifier only */
  int32_t input_grads_only; /* backward: skip every weight-gradient GEMM (attribution wants dvideo/daudio only); grads
                             * must still point at a scratch buffer of n_params floats for the small reductions */
  const void* fused_in;  /* stage 2: classifier input [B,fused] dtype */
  const void* dfused_in; /* stage 1 backward: gradient of the fused embedding [B,fused] dtype */
  void* dfused_out;      /* stage 2 backward: optional gradient w.r.t. fused_in */
  /* Data-parallel overlap (stage 0 backward only): grad_events[k] (cudaEvent_t) is recorded on the stream as soon as
   * gradient bucket k of the flat buffer is final, in the order  classifier + out_norm | layer L-1 | ... | layer 0 |
   * projections, input norms and pos_embed.  n_grad_events must be 0 or layers + 2.  The caller waits for event k on a
   * communication stream and all-reduces bucket k while backward continues. */
  void* const* grad_events;
  int32_t n_grad_events;
  /* SyncBatchNorm for the train.py variant under data parallelism (train.py:51-52,66-74,116,125 compute batch statistics
   * over the WHOLE batch): bn_world > 1 makes every BatchNorm layer call bn_sync(bn_sync_user, buf, n, stream) on the
   * host, while it enqueues its kernels, for each buffer of partial column sums; the callee must enqueue an in-place
   * SUM all-reduce of buf[0..n) over the bn_world replicas on `stream` (ordered after what is already enqueued) and
   * return 0.  Forward: sum(x), then sum((x - mean)^2) (two calls of C floats per layer); backward: [sum(dy),
   * sum(dy * xhat)] (one call of 2C floats).  Row counts are multiplied by bn_world (equal shards).  bn_world <= 1 or
   * bn_sync NULL: plain per-replica BatchNorm. */
  int32_t bn_world;
  int (*bn_sync)(void* user, float* buf, int64_t n, void* stream);
  void* bn_sync_user;
  /* variant 2 with `use_layernorm=False` (train2.py:96,104-105,121 and :208,215; back-end/app/libs/model.py:16,23-24,39,
   * 88,95): bit MMER_NORM_FUSION_IDENTITY -- norm_video, norm_audio and out_norm are nn.Identity (stage 0 / 1);
   * bit MMER_NORM_HEAD_BATCHNORM -- the classifier normalises with nn.BatchNorm1d instead of nn.LayerNorm (stage 0 / 2):
   * C1 / C5 are the BatchNorm weights and biases and bn_state = [running_mean, running_var] of net[1], then of net[5]
   * (4 * hidden floats).  0 = the model as every caller in the reference builds it. */
  int32_t norms;
} mmer_model;
#define MMER_NORM_FUSION_IDENTITY 1
#define MMER_NORM_HEAD_BATCHNORM 2

/* Integrated Gradients around the model (captum.attr.IntegratedGradients.attribute as called at train2.py:826-834 and
 * back-end/app/libs/inference.py:313-321; Captum's default method "gausslegendre", multiply_by_inputs = True).
 * expand: out[k*n + i] = base[i] + alphas[k] * (x[i] - base[i])   (base NULL = zeros; step-major like Captum's cat)
 * reduce: attr[i] = (x[i] - base[i]) * sum_k weights[k] * grads[k*n + i]            (attr fp32)
 * n = n_per_step elements of one un-expanded tensor (multiple of 8); alphas / weights fp32 [n_steps] on the device. */
int mmer_ig_expand(const void* x, const void* base, const float* alphas, void* out, int64_t n_per_step, int64_t n_steps,
                   int in_dtype, int out_dtype, void* stream);
int mmer_ig_reduce(const void* grads, const void* x, const void* base, const float* weights, float* attr,
                   int64_t n_per_step, int64_t n_steps, int in_dtype, int grad_dtype, void* stream);

/* Batch assembly on the device, the step before the model in the reference's loop (train2.py:362-463):
 * mmer_feature_stats: per-feature mean and unbiased std + eps over x[R, D] (train2.py:430-441: `all_video.mean(dim=0)`,
 *   `all_video.std(dim=0) + 1e-6`); two passes, double accumulation; scratch = 2*D doubles.
 * mmer_collate: the closure collate_fn (train2.py:418-440) over a feature set resident in HBM -- frames[total, Dv] with
 *   offsets[N+1] (frame range of sample i), audio[N, Da], labels[N]; idx[B] selects the batch.  Writes
 *   video_out[B, Tmax, Dv] = pad_sequence(..., padding_value=0) of (x - mean_v) / std_v (train2.py:443-447; mean/std
 *   NULL = features already normalised), audio_out[B, Da], labels_out[B] and mask_out[B, Tmax] (1 = padded).
 *   Tmax must be >= the longest selected sample (the caller knows the lengths); out_dtype MMER_F32 or MMER_BF16. */
int mmer_feature_stats(const float* x, int64_t R, int64_t D, float eps, float* mean, float* std, double* scratch,
                       void* stream);
/* bf16-resident variant: mmer_normalize_rows z-scores (mean/std NULL = plain cast) and rounds x[R, D] to bf16 once;
 * mmer_collate_bf16 then gathers / pads / masks 2-byte rows (same bits as mmer_collate with bf16 output). */
int mmer_normalize_rows(const float* x, const float* mean, const float* std, void* out_bf16, int64_t R, int64_t D,
                        void* stream);
int mmer_collate_bf16(const void* frames_bf16, const int64_t* offsets, const void* audio_bf16, const int64_t* labels,
                      const int64_t* idx, void* video_out, void* audio_out, int64_t* labels_out, uint8_t* mask_out,
                      int64_t B, int64_t Tmax, int64_t Dv, int64_t Da, void* stream);
int mmer_collate(const float* frames, const int64_t* offsets, const float* audio, const int64_t* labels, const int64_t* idx,
                 const float* mean_v, const float* std_v, const float* mean_a, const float* std_a, void* video_out,
                 void* audio_out, int64_t* labels_out, uint8_t* mask_out, int64_t B, int64_t Tmax, int64_t Dv, int64_t Da,
                 int out_dtype, void* stream);

/* Evaluation bookkeeping of the validation / test loops (train2.py:593-607, 651-667, 724-741) without host round
 * trips: predicted[b] = argmax_c probs[b, c] (first maximum, torch.max) and confusion[y * C + predicted] += 1 (int64,
 * rows = true class like sklearn.metrics.confusion_matrix; the caller zeroes it at the start of an epoch).  predicted
 * may be NULL.  Accuracy and macro / micro precision / recall / F1 follow from the matrix. */
int mmer_eval_accumulate(const float* probs, const int64_t* labels, int64_t* predicted, int64_t* confusion, int64_t B,
                         int64_t C, void* stream);

/* Plain event plumbing for callers that only hold raw stream handles (the events above). */
int mmer_event_create(void** event_out);
int mmer_event_destroy(void* event);
int mmer_stream_wait_event(void* stream, void* event);

/* Batch-1 serving forward as ONE launch (the live request of back-end/app/libs/inference.py:494-495: one clip window of
 * <= 5 chunks): a single thread-block cluster walks the whole train2.py model in eval mode; weights stream once, split
 * by output feature over the cluster's CTAs; activations are exchanged through `scratch` (mmer_serve_scratch_bytes(),
 * fp32, stays L2-resident) between cluster barriers.  m: variant 2, dtype MMER_BF16, B == 1, T + 1 <= 16, fused 512,
 * 8 heads; uses params, shadow, off_g / off_l, video [T, video_dim] and audio [audio_dim] (bf16), mask, logits, probs.
 * Anything else (batches, longer clips, attention weights) goes through mmer_model_forward.
 * Up to 8 tokens at the train2.py default widths a second kernel (csrc/serve_small.cu) needs two cluster barriers per
 * layer instead of four: attention is computed where the head's q / k / v are produced, out_proj and linear2 are
 * split-K partial products summed by fp32 reductions in L2 (so the last bits may differ between identical calls;
 * MMER_DEBUG_SERVE_GLOBAL = 1 selects the bit-reproducible kernel).  It reads the weights from `packed`, a copy of the
 * bf16 shadow in MMA-fragment order at the same offsets: mmer_serve_pack(m, packed, stream) fills it (same size as the
 * shadow; call again whenever the shadow is re-cast).  packed may be NULL: the first kernel is used. */
int64_t mmer_serve_scratch_bytes(void);
int mmer_serve_pack(const mmer_model* m, void* packed, void* stream);
int mmer_serve_forward(const mmer_model* m, void* scratch, const void* packed, void* stream);

int64_t mmer_workspace_bytes(const mmer_model* m);
int mmer_model_forward(const mmer_model* m, void* stream);
int mmer_model_backward(const mmer_model* m, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMER_H_ */
