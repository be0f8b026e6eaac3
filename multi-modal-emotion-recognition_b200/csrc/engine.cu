// Whole-model orchestration: MultimodalEmotionModel forward and backward as one C call each.
// Everything is enqueued asynchronously on the caller's stream; the workspace (activations
// saved for backward + gradient scratch) is a single caller-owned allocation carved here.
//
// Reference data flow: train2.py:128-193,235-238,281-292 (LayerNorm variant, `variant == 2`)
// and train.py:64-106,123-130,139-142 (BatchNorm variant, `variant == 1`).  The encoder layer is
// the post-norm nn.TransformerEncoderLayer (ReLU, batch_first=False) configured at
// train2.py:111-118 / train.py:54-57; tokens are kept batch-major here, which removes both
// permutes of the reference (train2.py:172,181).
#include "common.cuh"

namespace mmer {

int gemm_tc(const mmer_gemm_args& a, cudaStream_t st);
int gemm_simt(const mmer_gemm_args& a, cudaStream_t st);
int bn_fwd_sync(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var, void* y,
                float* stats_out, int64_t N, int64_t C, int dtype, int training, int relu, float momentum, float drop_p,
                uint64_t seed, uint32_t site, cudaStream_t st, const BnSync* sy);
int bn_bwd_sync(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, void* dx,
                float* dgamma, float* dbeta, float* scratch, int64_t N, int64_t C, int dtype, int training, int relu,
                float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, const BnSync* sy);
int mha_fwd_ex(const void* qkv, const uint8_t* mask, void* out, float* probs, int64_t B, int64_t T, int64_t H, int64_t d,
               int dtype, float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, float* lse);
int mha_bwd_ex(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias_qkv, int64_t B, int64_t T,
               int64_t H, int64_t d, int dtype, float drop_p, uint64_t seed, uint32_t site, cudaStream_t st, const float* lse,
               const void* fwd_out);
int gemm_ln_fwd(const void* A, const void* W, const float* bias, const void* residual, const float* gamma, const float* beta,
                void* z_out, void* y_out, float* stats, long long M, long long K, DropCfg drop, cudaStream_t st);
extern int g_debug[16];

// Post-norm sub-layer tails (out_proj -> norm1, linear2 -> norm2), three ways (bf16, d_model = 512, M >= 256):
//   0  GEMM -> a (sub-layer output), add_ln_fwd(x, a) -> y; backward re-reads x and a            (fp32 mode, small shapes)
//   1  ONE tcgen05 kernel (gemm_ln.cu): GEMM + bias + dropout + residual + LayerNorm -> z, y, stats
//   2  GEMM with the bias + dropout + residual epilogue -> z, add_ln_fwd(z) -> y (reads ONE tensor)
// Modes 1 and 2 store the pre-LayerNorm sum z where mode 0 stores the sub-layer output, and backward reads it through
// mmer_add_ln_bwd_z (4 tensor passes instead of 5).  Forward and backward take the decision from the same (dtype, dims,
// debug knob), so a forward / backward pair always agrees.  Measured on B200 (profiles/r02_gemm_ln_*.txt): the fully
// fused kernel is correct but not faster than mode 2 -- its epilogue traffic (residual tile in, z tile out, z back from
// L2) goes through the same shared-memory ports that feed the MMAs -- so mode 2 is the default.
static int ln_mode(const mmer_model* m) {
  if (!(m->dtype == MMER_BF16 && m->fused == 512 && m->ffn % 64 == 0 && (int64_t)m->B * (m->T + 1) >= 256)) return 0;
  const int k = g_debug[MMER_DEBUG_NO_LN_FUSE];   // 0 default, 1 force mode 0, 2 force the fully fused kernel
  return k == 1 ? 0 : (k == 2 ? 1 : 2);
}

// z[M,N] = residual + dropout(x[M,K] W[N,K]^T + b)
static int lin_fwd_res(const mmer_model* m, const void* x, int64_t M, int64_t K, int64_t offW, int64_t offB, const void* residual,
                       void* z, int64_t N, float drop_p, uint32_t site, cudaStream_t st);

bool gemm_tc_stages_output(const mmer_gemm_args& a);

int gemm_dispatch(const mmer_gemm_args& a, cudaStream_t st) {
  if (a.in_dtype == MMER_BF16) {
    MMER_TRY(gemm_tc(a, st));
    // column sums of D: free in the staged (TMA-store) epilogue, a separate pass otherwise
    if (a.d_colsum != nullptr && !gemm_tc_stages_output(a)) return mmer_colsum(a.D, a.d_colsum, a.M, a.N, a.ldd, a.out_dtype, st);
    return 0;
  }
  MMER_TRY(gemm_simt(a, st));
  if (a.d_colsum != nullptr) MMER_TRY(mmer_colsum(a.D, a.d_colsum, a.M, a.N, a.ldd, a.out_dtype, st));
  if (a.a_rowsum != nullptr) {
    // fp32 parity mode: the row sums of an MN-major A are the column sums of A as stored ([K][M])
    MMER_CHECK_ARG(a.a_major == MMER_MAJOR_MN, "gemm: a_rowsum needs an MN-major A");
    return mmer_colsum(a.A, a.a_rowsum, a.K, a.M, a.lda, a.in_dtype, st);
  }
  return 0;
}

struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)), off(0) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};

struct LayerWs {
  void *qkv, *att, *ao, *x1, *h, *f2, *x2;
  uint8_t* hmask;   // bf16 mode: 1 bit per element of h (stored value > 0): the ReLU+dropout gate of backward
  float *st1, *st2;
  float* lse;       // long-sequence attention: (max, 1 / sum) of every softmax row, forward -> backward
};
struct Ws {
  void *pv, *pa, *pvn, *pan, *x0;
  float *st_e, *st_bnv, *st_bna, *st_fc, *st_fc2;
  LayerWs L[MMER_MAX_LAYERS];
  float* pooled;
  void* fused;
  float* st_o;
  void *h1p, *h1, *h2p, *h2;
  float *st_h1, *st_h2;
  // backward scratch
  void *g_x, *g_z2, *g_f2, *g_h, *g_x1, *g_z1, *g_ao, *g_att, *g_qkv, *g_pv, *g_pa, *g_pvn, *g_pan;
  void *g_fused, *g_h1, *g_h1p, *g_h2, *g_h2p;
  float* bn_scratch;
  size_t total;
};

static void carve(const mmer_model* m, void* base, Ws* w) {
  Carver c(base);
  const size_t e = m->dtype == MMER_BF16 ? 2 : 4;
  const size_t B = m->B, T = m->T, S = T + 1, F = m->fused, Hd = m->hidden, FF = m->ffn;
  const size_t M = B * S, Mv = B * T;
  w->pv = c.take(Mv * F * e);
  w->pa = c.take(B * F * e);
  w->pvn = m->variant == 1 ? c.take(Mv * F * e) : nullptr;
  w->pan = m->variant == 1 ? c.take(B * F * e) : nullptr;
  w->x0 = c.take(M * F * e);
  w->st_e = (float*)c.take(M * 2 * 4);
  w->st_bnv = (float*)c.take(2 * F * 4);
  w->st_bna = (float*)c.take(2 * F * 4);
  w->st_fc = (float*)c.take(2 * Hd * 4);
  w->st_fc2 = (float*)c.take(2 * Hd * 4);
  for (int l = 0; l < m->layers; ++l) {
    LayerWs& L = w->L[l];
    L.qkv = c.take(M * 3 * F * e);
    L.att = c.take(M * F * e);
    L.ao = c.take(M * F * e);
    L.x1 = c.take(M * F * e);
    L.h = c.take(M * FF * e);
    L.hmask = (m->dtype == MMER_BF16 && FF % 64 == 0) ? (uint8_t*)c.take(M * FF / 8) : nullptr;
    L.f2 = c.take(M * F * e);
    L.x2 = c.take(M * F * e);
    L.st1 = (float*)c.take(M * 2 * 4);
    L.st2 = (float*)c.take(M * 2 * 4);
    // (max, 1 / sum) per softmax row, then the dropout keep bits of the row (one word per 32 keys): mha_long_stats_floats
    L.lse = S > 32 ? (float*)c.take(B * (size_t)m->heads * S * (2 + (S + 31) / 32) * 4) : nullptr;
  }
  w->pooled = (float*)c.take(B * F * 4);
  w->fused = c.take(B * F * e);
  w->st_o = (float*)c.take(B * 2 * 4);
  w->h1p = c.take(B * Hd * e);
  w->h1 = c.take(B * Hd * e);
  w->h2p = c.take(B * Hd * e);
  w->h2 = c.take(B * Hd * e);
  w->st_h1 = (float*)c.take(B * 2 * 4);
  w->st_h2 = (float*)c.take(B * 2 * 4);
  w->g_x = c.take(M * F * e);
  w->g_z2 = c.take(M * F * e);
  w->g_f2 = c.take(M * F * e);
  w->g_h = c.take(M * FF * e);
  w->g_x1 = c.take(M * F * e);
  w->g_z1 = c.take(M * F * e);
  w->g_ao = c.take(M * F * e);
  w->g_att = c.take(M * F * e);
  w->g_qkv = c.take(M * 3 * F * e);
  w->g_pv = c.take(Mv * F * e);
  w->g_pa = c.take(B * F * e);
  w->g_pvn = m->variant == 1 ? c.take(Mv * F * e) : nullptr;
  w->g_pan = m->variant == 1 ? c.take(B * F * e) : nullptr;
  w->g_fused = c.take(B * F * e);
  w->g_h1 = c.take(B * Hd * e);
  w->g_h1p = c.take(B * Hd * e);
  w->g_h2 = c.take(B * Hd * e);
  w->g_h2p = c.take(B * Hd * e);
  w->bn_scratch = (float*)c.take(2 * (F > Hd ? F : Hd) * 4);
  w->total = c.off;
}

static int validate(const mmer_model* m, bool need_ws) {
  MMER_CHECK_ARG(m != nullptr, "model: null");
  MMER_CHECK_ARG(m->variant == 1 || m->variant == 2, "model: variant must be 1 (train.py) or 2 (train2.py)");
  MMER_CHECK_ARG((m->norms & ~3) == 0 && (m->variant == 2 || m->norms == 0), "model: norms flags are for variant 2 only");
  MMER_CHECK_ARG(m->dtype == MMER_F32 || m->dtype == MMER_BF16, "model: bad dtype");
  MMER_CHECK_ARG(m->B > 0 && m->T > 0, "model: empty batch (B=%d T=%d)", m->B, m->T);
  MMER_CHECK_ARG(m->layers >= 1 && m->layers <= MMER_MAX_LAYERS, "model: layers out of range");
  MMER_CHECK_ARG(m->heads >= 1 && m->fused % m->heads == 0, "model: fused_dim not divisible by heads");
  const int d = m->fused / m->heads;
  MMER_CHECK_ARG(d == 32 || d == 64, "model: head dim %d unsupported (32 or 64)", d);
  MMER_CHECK_ARG(m->fused % 8 == 0 && m->hidden % 8 == 0 && m->ffn % 8 == 0 && m->video_dim % 8 == 0 &&
                     m->audio_dim % 8 == 0,
                 "model: every feature dimension must be a multiple of 8");
  MMER_CHECK_ARG(m->fused <= 2048 && m->hidden <= 2048, "model: fused/hidden width above 2048 unsupported");
  MMER_CHECK_ARG(m->classes >= 1 && m->classes <= 16, "model: classes must be <= 16");
  if (need_ws) {
    MMER_CHECK_ARG(m->params != nullptr, "model: params is null");
    MMER_CHECK_ARG(m->dtype != MMER_BF16 || m->shadow != nullptr, "model: bf16 mode needs the bf16 shadow weights");
    MMER_CHECK_ARG(m->variant != 1 || m->bn_state != nullptr, "model: variant 1 needs bn_state");
    MMER_CHECK_ARG(!(m->variant == 2 && (m->norms & MMER_NORM_HEAD_BATCHNORM)) || m->bn_state != nullptr,
                   "model: a BatchNorm classifier head needs bn_state");
    Ws w;
    carve(m, nullptr, &w);
    MMER_CHECK_ARG(m->workspace != nullptr && (size_t)m->workspace_bytes >= w.total,
                   "model: workspace too small (%lld < %lld)", (long long)m->workspace_bytes, (long long)w.total);
  }
  return 0;
}

static inline const float* P(const mmer_model* m, int64_t off) { return off < 0 ? nullptr : m->params + off; }
static inline float* G(const mmer_model* m, int64_t off) { return off < 0 ? nullptr : m->grads + off; }
static inline const void* Wt(const mmer_model* m, int64_t off) {
  return m->dtype == MMER_BF16 ? (const void*)(reinterpret_cast<const bf16*>(m->shadow) + off)
                               : (const void*)(m->params + off);
}

// y[M,N] = x[M,K] W[N,K]^T + b, optional relu / dropout
static int lin_fwd(const mmer_model* m, const void* x, int64_t M, int64_t K, int64_t offW, int64_t offB, void* y,
                   int64_t N, int relu, float drop_p, uint32_t site, cudaStream_t st, uint8_t* mask_out = nullptr) {
  mmer_gemm_args a = {};
  a.relu_mask_out = mask_out;
  a.A = x; a.B = Wt(m, offW); a.D = y; a.bias = P(m, offB);
  a.M = M; a.N = N; a.K = K; a.lda = K; a.ldb = K; a.ldd = N;
  a.a_major = MMER_MAJOR_K; a.b_major = MMER_MAJOR_K;
  a.in_dtype = m->dtype; a.out_dtype = m->dtype; a.relu = relu;
  a.drop_p = drop_p; a.seed = m->seed; a.drop_site = site;
  return gemm_dispatch(a, st);
}
static int lin_fwd_res(const mmer_model* m, const void* x, int64_t M, int64_t K, int64_t offW, int64_t offB, const void* residual,
                       void* z, int64_t N, float drop_p, uint32_t site, cudaStream_t st) {
  mmer_gemm_args a = {};
  a.A = x; a.B = Wt(m, offW); a.D = z; a.bias = P(m, offB); a.residual = residual;
  a.M = M; a.N = N; a.K = K; a.lda = K; a.ldb = K; a.ldd = N;
  a.a_major = MMER_MAJOR_K; a.b_major = MMER_MAJOR_K;
  a.in_dtype = m->dtype; a.out_dtype = m->dtype;
  a.drop_p = drop_p; a.seed = m->seed; a.drop_site = site;
  return gemm_dispatch(a, st);
}
// dx[M,K] = dy[M,N] W[N,K] (+ residual) (* gate)
static int lin_dgrad(const mmer_model* m, const void* dy, int64_t M, int64_t N, int64_t offW, int64_t K, void* dx,
                     const void* residual, const void* gate, float gate_scale, cudaStream_t st,
                     const uint8_t* gate_bits = nullptr, float* dx_colsum = nullptr) {
  mmer_gemm_args a = {};
  a.gate_bits = gate_bits;
  a.d_colsum = m->input_grads_only ? nullptr : dx_colsum;
  a.A = dy; a.B = Wt(m, offW); a.D = dx; a.residual = residual; a.gate = gate; a.gate_scale = gate_scale;
  a.M = M; a.N = K; a.K = N; a.lda = N; a.ldb = K; a.ldd = K;
  a.a_major = MMER_MAJOR_K; a.b_major = MMER_MAJOR_MN;
  a.in_dtype = m->dtype; a.out_dtype = m->dtype;
  return gemm_dispatch(a, st);
}
// gW[N,K] += dy[M,N]^T x[M,K];  gb[N] += column sums of dy (offB < 0: no bias gradient wanted)
static int lin_wgrad(const mmer_model* m, const void* dy, const void* x, int64_t M, int64_t N, int64_t K,
                     int64_t offW, int64_t offB, cudaStream_t st) {
  if (m->input_grads_only) return 0;   // attribution: only the data gradients are wanted
  mmer_gemm_args a = {};
  a.A = dy; a.B = x; a.D = G(m, offW); a.a_rowsum = G(m, offB);
  a.M = N; a.N = K; a.K = M; a.lda = N; a.ldb = K; a.ldd = K;
  a.a_major = MMER_MAJOR_MN; a.b_major = MMER_MAJOR_MN;
  a.in_dtype = m->dtype; a.out_dtype = MMER_F32; a.accumulate = 1;
  return gemm_dispatch(a, st);
}

static inline uint32_t site_layer(int l, int k) { return 10u + 4u * (uint32_t)l + (uint32_t)k; }

struct Dims {
  int64_t B, T, S, F, Hd, FF, M, Mv;
  int dt;
  bool tr;
  float pf, pc;
  uint64_t seed;
  const uint8_t* mask;
  const int64_t* g;
  BnSync sy;
  bool ln_fusion, ln_head, bn_head;   // which normalisation the fusion module / the classifier head use
  explicit Dims(const mmer_model* m)
      : B(m->B), T(m->T), S(m->T + 1), F(m->fused), Hd(m->hidden), FF(m->ffn), M((int64_t)m->B * (m->T + 1)),
        Mv((int64_t)m->B * m->T), dt(m->dtype), tr(m->training != 0), pf(tr ? m->p_fusion : 0.f),
        pc(tr ? m->p_classifier : 0.f), seed(m->seed), mask(m->has_mask ? m->mask : nullptr), g(m->off_g),
        sy{m->bn_sync, m->bn_sync_user, m->bn_world},
        ln_fusion(m->variant == 2 && !(m->norms & MMER_NORM_FUSION_IDENTITY)),
        ln_head(m->variant == 2 && !(m->norms & MMER_NORM_HEAD_BATCHNORM)),
        bn_head(m->variant == 2 && (m->norms & MMER_NORM_HEAD_BATCHNORM)) {}
};

// CrossModalFusion.forward: projections -> token assembly -> encoder layers -> pooling (+ out_norm)
static int fusion_forward(const mmer_model* m, Ws& w, cudaStream_t st) {
  const Dims d(m);
  const int64_t B = d.B, T = d.T, F = d.F, M = d.M, Mv = d.Mv, FF = d.FF;
  const int64_t* g = d.g;
  MMER_TRY(lin_fwd(m, m->video, Mv, m->video_dim, g[MMER_G_WV], g[MMER_G_BV], w.pv, F, 0, 0.f, 0, st));
  MMER_TRY(lin_fwd(m, m->audio, B, m->audio_dim, g[MMER_G_WA], g[MMER_G_BA], w.pa, F, 0, 0.f, 0, st));
  if (d.ln_fusion) {
    MMER_TRY(mmer_embed_fwd(w.pv, w.pa, P(m, g[MMER_G_NV_W]), P(m, g[MMER_G_NV_B]), P(m, g[MMER_G_NA_W]),
                            P(m, g[MMER_G_NA_B]), P(m, g[MMER_G_POS]), w.x0, w.st_e, B, T, F, d.dt, d.pf, d.seed, 0, st));
  } else if (m->variant == 2) {   // use_layernorm=False: norm_video / norm_audio are nn.Identity (train2.py:104-105)
    MMER_TRY(mmer_embed_fwd(w.pv, w.pa, nullptr, nullptr, nullptr, nullptr, P(m, g[MMER_G_POS]), w.x0, w.st_e, B, T, F,
                            d.dt, d.pf, d.seed, 0, st));
  } else {
    float* bs = m->bn_state;
    MMER_TRY(bn_fwd_sync(w.pv, P(m, g[MMER_G_NV_W]), P(m, g[MMER_G_NV_B]), bs, bs + F, w.pvn, w.st_bnv, Mv, F, d.dt, d.tr,
                         0, 0.1f, 0.f, d.seed, 0, st, &d.sy));
    MMER_TRY(bn_fwd_sync(w.pa, P(m, g[MMER_G_NA_W]), P(m, g[MMER_G_NA_B]), bs + 2 * F, bs + 3 * F, w.pan, w.st_bna, B, F,
                         d.dt, d.tr, 0, 0.1f, 0.f, d.seed, 0, st, &d.sy));
    MMER_TRY(mmer_embed_fwd(w.pvn, w.pan, nullptr, nullptr, nullptr, nullptr, P(m, g[MMER_G_POS]), w.x0, w.st_e, B, T, F,
                            d.dt, 0.f, d.seed, 0, st));
  }
  const void* x = w.x0;
  const int64_t SS = d.S * d.S;
  for (int l = 0; l < m->layers; ++l) {
    const int64_t* o = m->off_l[l];
    LayerWs& L = w.L[l];
    MMER_TRY(lin_fwd(m, x, M, F, o[MMER_L_IN_W], o[MMER_L_IN_B], L.qkv, 3 * F, 0, 0.f, 0, st));
    float* probs = m->attn_probs ? m->attn_probs + (int64_t)l * B * m->heads * SS : nullptr;
    MMER_TRY(mha_fwd_ex(L.qkv, d.mask, L.att, probs, B, T, m->heads, F / m->heads, d.dt, d.pf, d.seed, site_layer(l, 0), st,
                        L.lse));
    const int lnm = ln_mode(m);   // modes 1 and 2: L.ao / L.f2 hold z1 / z2 (pre-LayerNorm sums)
    if (lnm == 1) {
      MMER_TRY(gemm_ln_fwd(L.att, Wt(m, o[MMER_L_OUT_W]), P(m, o[MMER_L_OUT_B]), x, P(m, o[MMER_L_N1_W]), P(m, o[MMER_L_N1_B]),
                           L.ao, L.x1, L.st1, M, F, make_drop(d.pf, d.seed, site_layer(l, 1)), st));
    } else if (lnm == 2) {
      MMER_TRY(lin_fwd_res(m, L.att, M, F, o[MMER_L_OUT_W], o[MMER_L_OUT_B], x, L.ao, F, d.pf, site_layer(l, 1), st));
      MMER_TRY(mmer_add_ln_fwd(nullptr, L.ao, P(m, o[MMER_L_N1_W]), P(m, o[MMER_L_N1_B]), L.x1, L.st1, M, F, d.dt, 0, 0.f, 0,
                               0.f, 0, d.seed, st));
    } else {
      MMER_TRY(lin_fwd(m, L.att, M, F, o[MMER_L_OUT_W], o[MMER_L_OUT_B], L.ao, F, 0, 0.f, 0, st));
      MMER_TRY(mmer_add_ln_fwd(x, L.ao, P(m, o[MMER_L_N1_W]), P(m, o[MMER_L_N1_B]), L.x1, L.st1, M, F, d.dt, 0, d.pf,
                               site_layer(l, 1), 0.f, 0, d.seed, st));
    }
    MMER_TRY(lin_fwd(m, L.x1, M, F, o[MMER_L_FF1_W], o[MMER_L_FF1_B], L.h, FF, 1, d.pf, site_layer(l, 2), st,
                     d.tr ? L.hmask : nullptr));
    if (lnm == 1) {
      MMER_TRY(gemm_ln_fwd(L.h, Wt(m, o[MMER_L_FF2_W]), P(m, o[MMER_L_FF2_B]), L.x1, P(m, o[MMER_L_N2_W]), P(m, o[MMER_L_N2_B]),
                           L.f2, L.x2, L.st2, M, FF, make_drop(d.pf, d.seed, site_layer(l, 3)), st));
    } else if (lnm == 2) {
      MMER_TRY(lin_fwd_res(m, L.h, M, FF, o[MMER_L_FF2_W], o[MMER_L_FF2_B], L.x1, L.f2, F, d.pf, site_layer(l, 3), st));
      MMER_TRY(mmer_add_ln_fwd(nullptr, L.f2, P(m, o[MMER_L_N2_W]), P(m, o[MMER_L_N2_B]), L.x2, L.st2, M, F, d.dt, 0, 0.f, 0,
                               0.f, 0, d.seed, st));
    } else {
      MMER_TRY(lin_fwd(m, L.h, M, FF, o[MMER_L_FF2_W], o[MMER_L_FF2_B], L.f2, F, 0, 0.f, 0, st));
      MMER_TRY(mmer_add_ln_fwd(L.x1, L.f2, P(m, o[MMER_L_N2_W]), P(m, o[MMER_L_N2_B]), L.x2, L.st2, M, F, d.dt, 0, d.pf,
                               site_layer(l, 3), 0.f, 0, d.seed, st));
    }
    x = L.x2;
  }
  MMER_TRY(mmer_pool_ln_fwd(x, d.mask, d.ln_fusion ? P(m, g[MMER_G_ON_W]) : nullptr,
                            d.ln_fusion ? P(m, g[MMER_G_ON_B]) : nullptr, w.pooled, w.fused, w.st_o, B, T, F, d.dt, st));
  if (m->fused_out) {
    cudaError_t e = cudaMemcpyAsync(m->fused_out, w.fused, (size_t)B * F * (d.dt == MMER_BF16 ? 2 : 4),
                                    cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "copy fused");
  }
  return 0;
}

// EmotionClassifier.forward (+ the softmax of MultimodalEmotionModel.forward)
static int head_forward(const mmer_model* m, Ws& w, const void* fused, cudaStream_t st) {
  const Dims d(m);
  const int64_t B = d.B, F = d.F, Hd = d.Hd;
  const int64_t* g = d.g;
  MMER_TRY(lin_fwd(m, fused, B, F, g[MMER_G_C0_W], g[MMER_G_C0_B], w.h1p, Hd, 0, 0.f, 0, st));
  if (d.bn_head) {   // use_layernorm=False: Linear -> BatchNorm1d -> ReLU -> Dropout, twice (train2.py:215-228)
    float* bs = m->bn_state;
    MMER_TRY(bn_fwd_sync(w.h1p, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), bs, bs + Hd, w.h1, w.st_fc, B, Hd, d.dt, d.tr,
                         1, 0.1f, d.pc, d.seed, 200, st, &d.sy));
    MMER_TRY(lin_fwd(m, w.h1, B, Hd, g[MMER_G_C4_W], g[MMER_G_C4_B], w.h2p, Hd, 0, 0.f, 0, st));
    MMER_TRY(bn_fwd_sync(w.h2p, P(m, g[MMER_G_C5_W]), P(m, g[MMER_G_C5_B]), bs + 2 * Hd, bs + 3 * Hd, w.h2, w.st_fc2, B, Hd,
                         d.dt, d.tr, 1, 0.1f, d.pc, d.seed, 201, st, &d.sy));
    MMER_TRY(mmer_head_out_fwd(w.h2, P(m, g[MMER_G_C8_W]), P(m, g[MMER_G_C8_B]), m->logits, m->probs, B, Hd, m->classes,
                               d.dt, st));
  } else if (m->variant == 2) {
    MMER_TRY(mmer_add_ln_fwd(nullptr, w.h1p, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), w.h1, w.st_h1, B, Hd, d.dt, 1,
                             0.f, 0, d.pc, 200, d.seed, st));
    MMER_TRY(lin_fwd(m, w.h1, B, Hd, g[MMER_G_C4_W], g[MMER_G_C4_B], w.h2p, Hd, 0, 0.f, 0, st));
    MMER_TRY(mmer_add_ln_fwd(nullptr, w.h2p, P(m, g[MMER_G_C5_W]), P(m, g[MMER_G_C5_B]), w.h2, w.st_h2, B, Hd, d.dt, 1,
                             0.f, 0, d.pc, 201, d.seed, st));
    MMER_TRY(mmer_head_out_fwd(w.h2, P(m, g[MMER_G_C8_W]), P(m, g[MMER_G_C8_B]), m->logits, m->probs, B, Hd, m->classes,
                               d.dt, st));
  } else {
    float* bs = m->bn_state + 4 * F;
    MMER_TRY(bn_fwd_sync(w.h1p, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), bs, bs + Hd, w.h1, w.st_fc, B, Hd, d.dt, d.tr,
                         1, 0.1f, d.pc, d.seed, 200, st, &d.sy));
    MMER_TRY(mmer_head_out_fwd(w.h1, P(m, g[MMER_G_C8_W]), P(m, g[MMER_G_C8_B]), m->logits, m->probs, B, Hd, m->classes,
                               d.dt, st));
  }
  return 0;
}

int model_forward(const mmer_model* m, cudaStream_t st) {
  MMER_TRY(validate(m, true));
  MMER_CHECK_ARG(m->stage >= 0 && m->stage <= 2, "model_forward: bad stage");
  MMER_CHECK_ARG(m->stage == 2 || (m->video && m->audio), "model_forward: video/audio must be set");
  MMER_CHECK_ARG(m->stage == 1 || m->logits, "model_forward: logits must be set");
  MMER_CHECK_ARG(m->stage != 2 || m->fused_in, "model_forward: stage 2 needs fused_in");
  MMER_CHECK_ARG(m->stage != 1 || m->fused_out, "model_forward: stage 1 needs fused_out");
  Ws w;
  carve(m, m->workspace, &w);
  if (m->stage != 2) MMER_TRY(fusion_forward(m, w, st));
  if (m->stage != 1) MMER_TRY(head_forward(m, w, m->stage == 2 ? m->fused_in : w.fused, st));
  return 0;
}

// gradient of the classifier head; leaves d(fused) in w.g_fused
static int head_backward(const mmer_model* m, Ws& w, const void* fused, cudaStream_t st) {
  const Dims d(m);
  const int64_t B = d.B, F = d.F, Hd = d.Hd;
  const int64_t* g = d.g;
  if (d.bn_head) {
    MMER_TRY(mmer_head_out_bwd(m->dlogits, w.h2, P(m, g[MMER_G_C8_W]), w.g_h2, G(m, g[MMER_G_C8_W]), G(m, g[MMER_G_C8_B]),
                               B, Hd, m->classes, d.dt, st));
    MMER_TRY(bn_bwd_sync(w.g_h2, w.h2p, w.st_fc2, P(m, g[MMER_G_C5_W]), P(m, g[MMER_G_C5_B]), w.g_h2p, G(m, g[MMER_G_C5_W]),
                         G(m, g[MMER_G_C5_B]), w.bn_scratch, B, Hd, d.dt, d.tr, 1, d.pc, d.seed, 201, st, &d.sy));
    MMER_TRY(lin_wgrad(m, w.g_h2p, w.h1, B, Hd, Hd, g[MMER_G_C4_W], g[MMER_G_C4_B], st));
    MMER_TRY(lin_dgrad(m, w.g_h2p, B, Hd, g[MMER_G_C4_W], Hd, w.g_h1, nullptr, nullptr, 0.f, st));
    MMER_TRY(bn_bwd_sync(w.g_h1, w.h1p, w.st_fc, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), w.g_h1p, G(m, g[MMER_G_C1_W]),
                         G(m, g[MMER_G_C1_B]), w.bn_scratch, B, Hd, d.dt, d.tr, 1, d.pc, d.seed, 200, st, &d.sy));
  } else if (m->variant == 2) {
    MMER_TRY(mmer_head_out_bwd(m->dlogits, w.h2, P(m, g[MMER_G_C8_W]), w.g_h2, G(m, g[MMER_G_C8_W]), G(m, g[MMER_G_C8_B]),
                               B, Hd, m->classes, d.dt, st));
    MMER_TRY(mmer_add_ln_bwd(w.g_h2, nullptr, w.h2p, w.st_h2, P(m, g[MMER_G_C5_W]), P(m, g[MMER_G_C5_B]), w.g_h2p, nullptr,
                             G(m, g[MMER_G_C5_W]), G(m, g[MMER_G_C5_B]), G(m, g[MMER_G_C4_B]), B, Hd, d.dt, 1, 0.f, 0, d.pc,
                             201, d.seed, st));
    MMER_TRY(lin_wgrad(m, w.g_h2p, w.h1, B, Hd, Hd, g[MMER_G_C4_W], -1, st));
    MMER_TRY(lin_dgrad(m, w.g_h2p, B, Hd, g[MMER_G_C4_W], Hd, w.g_h1, nullptr, nullptr, 0.f, st));
    MMER_TRY(mmer_add_ln_bwd(w.g_h1, nullptr, w.h1p, w.st_h1, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), w.g_h1p, nullptr,
                             G(m, g[MMER_G_C1_W]), G(m, g[MMER_G_C1_B]), G(m, g[MMER_G_C0_B]), B, Hd, d.dt, 1, 0.f, 0, d.pc,
                             200, d.seed, st));
  } else {
    MMER_TRY(mmer_head_out_bwd(m->dlogits, w.h1, P(m, g[MMER_G_C8_W]), w.g_h1, G(m, g[MMER_G_C8_W]), G(m, g[MMER_G_C8_B]),
                               B, Hd, m->classes, d.dt, st));
    MMER_TRY(bn_bwd_sync(w.g_h1, w.h1p, w.st_fc, P(m, g[MMER_G_C1_W]), P(m, g[MMER_G_C1_B]), w.g_h1p, G(m, g[MMER_G_C1_W]),
                         G(m, g[MMER_G_C1_B]), w.bn_scratch, B, Hd, d.dt, d.tr, 1, d.pc, d.seed, 200, st, &d.sy));
  }
  MMER_TRY(lin_wgrad(m, w.g_h1p, fused, B, Hd, F, g[MMER_G_C0_W], d.ln_head ? -1 : g[MMER_G_C0_B], st));
  MMER_TRY(lin_dgrad(m, w.g_h1p, B, Hd, g[MMER_G_C0_W], F, w.g_fused, nullptr, nullptr, 0.f, st));
  return 0;
}

// gradient bucket k of the flat buffer is final: tell the data-parallel caller (see mmer_model.grad_events)
static int bucket_done(const mmer_model* m, int k, cudaStream_t st) {
  if (m->stage != 0 || m->n_grad_events == 0) return 0;
  cudaError_t e = cudaEventRecord((cudaEvent_t)m->grad_events[k], st);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaEventRecord(grad bucket)");
}

static int fusion_backward(const mmer_model* m, Ws& w, const void* dfused, cudaStream_t st) {
  const Dims d(m);
  const int64_t B = d.B, T = d.T, F = d.F, M = d.M, Mv = d.Mv, FF = d.FF;
  const int64_t* g = d.g;
  const float pf = d.pf;
  MMER_TRY(mmer_pool_ln_bwd(dfused, w.pooled, w.st_o, d.ln_fusion ? P(m, g[MMER_G_ON_W]) : nullptr, d.mask, w.g_x,
                            d.ln_fusion ? G(m, g[MMER_G_ON_W]) : nullptr,
                            d.ln_fusion ? G(m, g[MMER_G_ON_B]) : nullptr, B, T, F, d.dt, st));
  MMER_TRY(bucket_done(m, 0, st));   // classifier + out_norm
  const float relu_gate_scale = pf > 0.f ? make_drop(pf, d.seed, 0).scale : 1.f;
  for (int l = m->layers - 1; l >= 0; --l) {
    const int64_t* o = m->off_l[l];
    LayerWs& L = w.L[l];
    const void* xin = l == 0 ? w.x0 : w.L[l - 1].x2;
    // norm2 <- linear2
    void* d_f2 = pf > 0.f ? w.g_f2 : w.g_z2;
    if (ln_mode(m) != 0) {
      MMER_TRY(mmer_add_ln_bwd_z(w.g_x, L.f2, L.st2, P(m, o[MMER_L_N2_W]), w.g_z2, pf > 0.f ? w.g_f2 : nullptr,
                                 G(m, o[MMER_L_N2_W]), G(m, o[MMER_L_N2_B]), G(m, o[MMER_L_FF2_B]), M, F, d.dt, pf,
                                 site_layer(l, 3), d.seed, st));
    } else {
      MMER_TRY(mmer_add_ln_bwd(w.g_x, L.x1, L.f2, L.st2, P(m, o[MMER_L_N2_W]), nullptr, w.g_z2, pf > 0.f ? w.g_f2 : nullptr,
                               G(m, o[MMER_L_N2_W]), G(m, o[MMER_L_N2_B]), G(m, o[MMER_L_FF2_B]), M, F, d.dt, 0, pf,
                               site_layer(l, 3), 0.f, 0, d.seed, st));
    }
    MMER_TRY(lin_wgrad(m, d_f2, L.h, M, F, FF, o[MMER_L_FF2_W], -1, st));
    // through ReLU (+ its dropout): gate on the stored post-activation.  No gradient tensor is re-read for a bias
    // gradient: add_ln_bwd, mha_bwd and embed_bwd sum the columns of what they store while it is on chip; for linear1
    // (and the BatchNorm variant's projections) the weight-gradient GEMM adds the row sums of its A operand (a_rowsum),
    // which costs that GEMM ~20 % and is therefore used only where no producer can do it
    // the gate is read as the bit mask the forward epilogue wrote (1/16 of re-reading h) when there is one
    const uint8_t* hmask = d.tr ? L.hmask : nullptr;   // written by the forward pass in training mode only
    // linear1's bias gradient = column sums of g_h: taken from the staged tile in this GEMM's epilogue (it used to be a
    // row-sum MMA inside linear1's weight-gradient GEMM, ~20 % of that kernel)
    MMER_TRY(lin_dgrad(m, d_f2, M, F, o[MMER_L_FF2_W], FF, w.g_h, nullptr, hmask ? nullptr : L.h, relu_gate_scale, st, hmask,
                       G(m, o[MMER_L_FF1_B])));
    MMER_TRY(lin_wgrad(m, w.g_h, L.x1, M, FF, F, o[MMER_L_FF1_W], -1, st));
    MMER_TRY(lin_dgrad(m, w.g_h, M, FF, o[MMER_L_FF1_W], F, w.g_x1, w.g_z2, nullptr, 0.f, st));
    // norm1 <- attention
    void* d_ao = pf > 0.f ? w.g_ao : w.g_z1;
    if (ln_mode(m) != 0) {
      MMER_TRY(mmer_add_ln_bwd_z(w.g_x1, L.ao, L.st1, P(m, o[MMER_L_N1_W]), w.g_z1, pf > 0.f ? w.g_ao : nullptr,
                                 G(m, o[MMER_L_N1_W]), G(m, o[MMER_L_N1_B]), G(m, o[MMER_L_OUT_B]), M, F, d.dt, pf,
                                 site_layer(l, 1), d.seed, st));
    } else {
      MMER_TRY(mmer_add_ln_bwd(w.g_x1, xin, L.ao, L.st1, P(m, o[MMER_L_N1_W]), nullptr, w.g_z1, pf > 0.f ? w.g_ao : nullptr,
                               G(m, o[MMER_L_N1_W]), G(m, o[MMER_L_N1_B]), G(m, o[MMER_L_OUT_B]), M, F, d.dt, 0, pf,
                               site_layer(l, 1), 0.f, 0, d.seed, st));
    }
    MMER_TRY(lin_wgrad(m, d_ao, L.att, M, F, F, o[MMER_L_OUT_W], -1, st));
    MMER_TRY(lin_dgrad(m, d_ao, M, F, o[MMER_L_OUT_W], F, w.g_att, nullptr, nullptr, 0.f, st));
    // in_proj bias gradient = column sums of g_qkv: the short-sequence attention kernels sum what they store; for long
    // sequences (bf16) the weight-gradient GEMM adds the row sums of its A operand instead of a separate pass over the
    // 1536-column gradient (cfg4: 66 us per layer against ~25 us inside the GEMM)
    const bool in_bias_by_wgrad = d.dt == MMER_BF16 && d.S > 32 && !m->input_grads_only;
    MMER_TRY(mha_bwd_ex(L.qkv, d.mask, w.g_att, w.g_qkv, in_bias_by_wgrad ? nullptr : G(m, o[MMER_L_IN_B]), B, T, m->heads,
                        F / m->heads, d.dt, pf, d.seed, site_layer(l, 0), st, L.lse, L.att));
    MMER_TRY(lin_wgrad(m, w.g_qkv, xin, M, 3 * F, F, o[MMER_L_IN_W], in_bias_by_wgrad ? o[MMER_L_IN_B] : -1, st));
    MMER_TRY(bucket_done(m, 1 + (m->layers - 1 - l), st));   // every gradient of layer l is final
    MMER_TRY(lin_dgrad(m, w.g_qkv, M, 3 * F, o[MMER_L_IN_W], F, w.g_x, w.g_z1, nullptr, 0.f, st));
  }
  const void* dpv = w.g_pv;
  const void* dpa = w.g_pa;
  if (m->variant == 2 && !d.ln_fusion) {   // Identity norms: the gradient rows pass through the dropout mask
    MMER_TRY(mmer_embed_bwd(w.g_x, w.pv, w.pa, w.st_e, nullptr, nullptr, w.g_pv, w.g_pa, nullptr, nullptr, nullptr, nullptr,
                            G(m, g[MMER_G_POS]), nullptr, nullptr, B, T, F, d.dt, pf, d.seed, 0, st));
  } else if (m->variant == 2) {
    MMER_TRY(mmer_embed_bwd(w.g_x, w.pv, w.pa, w.st_e, P(m, g[MMER_G_NV_W]), P(m, g[MMER_G_NA_W]), w.g_pv, w.g_pa,
                            G(m, g[MMER_G_NV_W]), G(m, g[MMER_G_NV_B]), G(m, g[MMER_G_NA_W]), G(m, g[MMER_G_NA_B]),
                            G(m, g[MMER_G_POS]), G(m, g[MMER_G_BV]), G(m, g[MMER_G_BA]), B, T, F, d.dt, pf, d.seed, 0, st));
  } else {
    MMER_TRY(mmer_embed_bwd(w.g_x, w.pvn, w.pan, w.st_e, nullptr, nullptr, w.g_pvn, w.g_pan, nullptr, nullptr, nullptr,
                            nullptr, G(m, g[MMER_G_POS]), nullptr, nullptr, B, T, F, d.dt, 0.f, d.seed, 0, st));
    MMER_TRY(bn_bwd_sync(w.g_pvn, w.pv, w.st_bnv, P(m, g[MMER_G_NV_W]), P(m, g[MMER_G_NV_B]), w.g_pv, G(m, g[MMER_G_NV_W]),
                         G(m, g[MMER_G_NV_B]), w.bn_scratch, Mv, F, d.dt, d.tr, 0, 0.f, d.seed, 0, st, &d.sy));
    MMER_TRY(bn_bwd_sync(w.g_pan, w.pa, w.st_bna, P(m, g[MMER_G_NA_W]), P(m, g[MMER_G_NA_B]), w.g_pa, G(m, g[MMER_G_NA_W]),
                         G(m, g[MMER_G_NA_B]), w.bn_scratch, B, F, d.dt, d.tr, 0, 0.f, d.seed, 0, st, &d.sy));
  }
  MMER_TRY(lin_wgrad(m, dpv, m->video, Mv, F, m->video_dim, g[MMER_G_WV], d.ln_fusion ? -1 : g[MMER_G_BV], st));
  MMER_TRY(lin_wgrad(m, dpa, m->audio, B, F, m->audio_dim, g[MMER_G_WA], d.ln_fusion ? -1 : g[MMER_G_BA], st));
  MMER_TRY(bucket_done(m, m->layers + 1, st));   // projections, input norms, pos_embed
  if (m->dvideo) MMER_TRY(lin_dgrad(m, dpv, Mv, F, g[MMER_G_WV], m->video_dim, m->dvideo, nullptr, nullptr, 0.f, st));
  if (m->daudio) MMER_TRY(lin_dgrad(m, dpa, B, F, g[MMER_G_WA], m->audio_dim, m->daudio, nullptr, nullptr, 0.f, st));
  return 0;
}

int model_backward(const mmer_model* m, cudaStream_t st) {
  MMER_TRY(validate(m, true));
  MMER_CHECK_ARG(m->stage >= 0 && m->stage <= 2, "model_backward: bad stage");
  MMER_CHECK_ARG(m->grads != nullptr, "model_backward: grads must be set");
  MMER_CHECK_ARG(m->stage == 1 || m->dlogits, "model_backward: dlogits must be set");
  MMER_CHECK_ARG(m->stage == 2 || (m->video && m->audio), "model_backward: video/audio must be set");
  MMER_CHECK_ARG(m->stage != 1 || m->dfused_in, "model_backward: stage 1 needs dfused_in");
  MMER_CHECK_ARG(m->stage != 2 || m->fused_in, "model_backward: stage 2 needs fused_in");
  MMER_CHECK_ARG(m->n_grad_events == 0 || (m->n_grad_events == m->layers + 2 && m->grad_events != nullptr),
                 "model_backward: n_grad_events must be 0 or layers + 2");
  Ws w;
  carve(m, m->workspace, &w);
  if (m->stage != 1) {
    MMER_TRY(head_backward(m, w, m->stage == 2 ? m->fused_in : w.fused, st));
    if (m->stage == 2) {
      if (m->dfused_out) {
        cudaError_t e = cudaMemcpyAsync(m->dfused_out, w.g_fused, (size_t)m->B * m->fused * (m->dtype == MMER_BF16 ? 2 : 4),
                                        cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_fail(e, "copy dfused");
      }
      return 0;
    }
  }
  return fusion_backward(m, w, m->stage == 1 ? m->dfused_in : w.g_fused, st);
}

}  // namespace mmer

using namespace mmer;

extern "C" {

int64_t mmer_workspace_bytes(const mmer_model* m) {
  if (validate(m, false) != 0) return -1;
  Ws w;
  carve(m, nullptr, &w);
  return (int64_t)w.total;
}
int mmer_model_forward(const mmer_model* m, void* stream) { return model_forward(m, (cudaStream_t)stream); }
int mmer_model_backward(const mmer_model* m, void* stream) { return model_backward(m, (cudaStream_t)stream); }

int mmer_gemm(const mmer_gemm_args* a, void* stream) {
  MMER_CHECK_ARG(a != nullptr && a->A && a->B && a->D, "gemm: null pointer");
  MMER_CHECK_ARG(a->in_dtype == MMER_F32 || a->in_dtype == MMER_BF16, "gemm: bad in_dtype");
  return gemm_dispatch(*a, (cudaStream_t)stream);
}

}  // extern "C"
