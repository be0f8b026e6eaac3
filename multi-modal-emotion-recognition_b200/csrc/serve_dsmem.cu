// Batch-1 serving forward, S = T + 1 <= 8 tokens (the live request: window_size 5, routers/infer.py:9), with the
// activations exchanged through DISTRIBUTED SHARED MEMORY instead of a global scratch buffer.
//
// Same walk as serve.cu (one 16-CTA cluster, weights split by output feature, skinny GEMMs on mma.sync with 16-byte weight
// loads into permuted A fragments).  What changes is how a CTA's slice of a layer output reaches the others: every output
// element is stored straight into the destination buffer of EVERY CTA of the cluster (st.shared::cluster through mapa),
// in the format its consumer reads -- q / k / v rows per head in bf16, the linear1 output as the bf16 input rows of
// linear2, sub-layer outputs in fp32 for the residual + LayerNorm -- so the phase that follows a cluster barrier starts
// from its own shared memory: no global round trip, no staging pass, and the barrier's release / acquire only has to
// cover shared-memory stores.  serve.cu measured >= 2.3 us per phase for the global version (barrier + L2 round trip,
// profiles/r02_serving_kernel_phases.txt); there are 11 barriers for two layers.
// Buffers that peers write are never the ones a CTA may still be reading in the same barrier interval (see the phase
// list in the kernel); head vectors get their own buffers for that reason.
#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

namespace {

constexpr int SD_THREADS = 512;
constexpr int SD_WARPS = 16;
constexpr int SD_ROWS = 8;               // token rows (S <= 8): one MMA n-tile
constexpr int SD_F = 512;
constexpr int SD_KX = 1024;              // widest input held in xs (audio_dim)
constexpr int SD_KH = 2048;              // linear2 input
constexpr int SD_LDX = SD_KX + 32;       // 64-byte skew: conflict-free 16-byte B-fragment loads
constexpr int SD_LDH = SD_KH + 32;
constexpr int SD_LDA = SD_F + 32;
constexpr float SD_EPS = 1e-5f;

struct ServeParamsD {
  int T, S;
  int video_dim, audio_dim, ffn, hidden, classes, layers, heads;
  const bf16* shadow;
  const float* params;
  int64_t off_g[MMER_G_COUNT];
  int64_t off_l[MMER_MAX_LAYERS][MMER_L_COUNT];
  const bf16* video;
  const bf16* audio;
  const uint8_t* mask;
  long long* stamps;                   // 64 x int64 (phase boundaries, CTA 0) or NULL
  float* logits;
  float* probs;
};

struct SmemD {
  bf16 xs[SD_ROWS * SD_LDX];           // local: GEMV input (video / audio rows, then the bf16 residual stream)
  bf16 hs[SD_ROWS * SD_LDH];           // REMOTE-written: relu(linear1) rows, input of linear2
  float xf[SD_ROWS * SD_F];            // local: residual stream, fp32
  float xadd[SD_ROWS * SD_F];          // REMOTE-written: projections / out_proj / linear2 outputs, fp32
  bf16 att[SD_ROWS * SD_LDA];          // local: attention output
  bf16 qkv[8][3 * SD_ROWS * 64];       // REMOTE-written: per head q, k, v rows
  float sc[8][SD_ROWS * SD_ROWS];      // local: scores / probabilities per head
  float part[SD_WARPS][8 * 20];        // local: partial D tiles of one round, [token][feature] (SD_PART_LD)
  float hbuf[2][SD_KH];                // REMOTE-written: head hidden vectors (pre-norm)
  float nrm[3][SD_KH];                 // local: head LayerNorm staging (values, gamma, beta)
  float red[64];
};

__device__ __forceinline__ uint32_t sd_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t sd_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sd_cluster_sync() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte stores into the same shared-memory variable of CTA `r` of the cluster
__device__ __forceinline__ void sd_st16(void* local, int r, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(r));
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sd_put8_f32(float* local, int r, const float (&v)[8]) {
  sd_st16(local, r, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  sd_st16(local + 4, r, __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
}
__device__ __forceinline__ void sd_put8_bf16(bf16* local, int r, const float (&v)[8]) {
  sd_st16(local, r, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

struct Stamper {
  long long* out;
  bool on, fine;
  int n;
  __device__ __forceinline__ void mark() {
    if (on && n < 63) {
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      out[1 + n] = tns;
      out[0] = n + 1;
    }
    ++n;
  }
  __device__ __forceinline__ void mark_fine() { if (fine) mark(); }
};

// The weight fragments of a warp's unit for the first MAXIT k-steps, loaded ahead of the cluster barrier that publishes
// the input rows (weights do not depend on activations): the L2 latency of a phase's first loads hides behind the
// previous phase's barrier.
template <int MAXIT>
struct WPre {
  uint4 a0[MAXIT], a1[MAXIT];
};
struct Plan {
  int tiles, ksplit, klen, units;
};
__device__ __forceinline__ Plan sd_plan(int K, int n0, int n1) {
  Plan pl;
  pl.tiles = (n1 - n0) >> 4;
  pl.ksplit = 1;
  while (pl.ksplit * 2 * pl.tiles <= SD_WARPS && (K / (pl.ksplit * 2)) % 32 == 0) pl.ksplit *= 2;
  pl.units = pl.tiles * pl.ksplit;
  pl.klen = K / pl.ksplit;
  return pl;
}
template <int MAXIT>
__device__ __forceinline__ void sd_prefetch(WPre<MAXIT>& w, const bf16* __restrict__ W, int K, int n0, int n1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const Plan pl = sd_plan(K, n0, n1);
  if (warp < pl.units) {
    const int tile = warp / pl.ksplit, ks = warp - tile * pl.ksplit;
    const bf16* w0 = W + (long long)(n0 + tile * 16 + g) * K + ks * pl.klen + t * 8;
    const bf16* w1 = w0 + (long long)8 * K;
#pragma unroll
    for (int i = 0; i < MAXIT; ++i)
      if (i * 32 < pl.klen) {
        w.a0[i] = __ldg(reinterpret_cast<const uint4*>(w0 + i * 32));
        w.a1[i] = __ldg(reinterpret_cast<const uint4*>(w1 + i * 32));
      }
  }
}

constexpr int SD_PART_LD = 20;   // floats per token row of a partial tile (16 features + pad: conflict-free fragment stores)

// emit8(r, s, n, v[8]) for every CTA r of the cluster, token s < S and 8-feature group n in [n0, n1):
// v = act(sum_k xs[s][k] W[n..n+7][k] + bias); all threads of the CTA call it.  `pre` holds the first k-steps of
// round 0 (sd_prefetch with the same W, K, n0, n1).
template <int MAXIT, typename Emit>
__device__ __forceinline__ void sd_linear(SmemD& sm, Stamper& stp, const WPre<MAXIT>& pre, const bf16* xs, int ldx, int K,
                                          const bf16* __restrict__ W, const float* __restrict__ bias, int n0, int n1, int S,
                                          bool relu, int nc, Emit emit8) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const Plan pl = sd_plan(K, n0, n1);
  for (int base = 0; base < pl.units; base += SD_WARPS) {
    const int u = base + warp;
    if (u < pl.units) {
      const int tile = u / pl.ksplit, ks = u - tile * pl.ksplit;
      const bf16* w0 = W + (long long)(n0 + tile * 16 + g) * K + ks * pl.klen + t * 8;
      const bf16* w1 = w0 + (long long)8 * K;
      const bf16* x0 = xs + g * ldx + ks * pl.klen + t * 8;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      int kb = 0;
      if (base == 0) {
#pragma unroll
        for (int i = 0; i < MAXIT; ++i)
          if (i * 32 < pl.klen) {
            const uint4 b = *reinterpret_cast<const uint4*>(x0 + i * 32);
            mma_bf16_16816(c, pre.a0[i].x, pre.a1[i].x, pre.a0[i].y, pre.a1[i].y, b.x, b.y);
            mma_bf16_16816(c, pre.a0[i].z, pre.a1[i].z, pre.a0[i].w, pre.a1[i].w, b.z, b.w);
          }
        kb = MAXIT * 32;
      }
#pragma unroll 8
      for (; kb < pl.klen; kb += 32) {
        const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(w0 + kb));
        const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(w1 + kb));
        const uint4 b = *reinterpret_cast<const uint4*>(x0 + kb);
        mma_bf16_16816(c, a0.x, a1.x, a0.y, a1.y, b.x, b.y);
        mma_bf16_16816(c, a0.z, a1.z, a0.w, a1.w, b.z, b.w);
      }
      float* pp = sm.part[warp];       // [token][feature], SD_PART_LD floats per token
      pp[(2 * t) * SD_PART_LD + g] = c[0];
      pp[(2 * t + 1) * SD_PART_LD + g] = c[1];
      pp[(2 * t) * SD_PART_LD + g + 8] = c[2];
      pp[(2 * t + 1) * SD_PART_LD + g + 8] = c[3];
    }
    stp.mark_fine();
    __syncthreads();
    stp.mark_fine();
    const int round_tiles = min(SD_WARPS, pl.units - base) / pl.ksplit;
    const int groups = round_tiles * 2;                // 8-feature groups of this round, contiguous in n
    const int tasks = groups * S * nc;                 // group fastest: lanes store runs of consecutive addresses
    for (int e = threadIdx.x; e < tasks; e += SD_THREADS) {
      const int grp = e % groups, rest = e / groups, s = rest % S, r = rest / S;
      const int tl = grp >> 1, half = grp & 1;
      float v[8];
      const int n = n0 + (base / pl.ksplit + tl) * 16 + half * 8;
      if (bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
        v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      for (int ks = 0; ks < pl.ksplit; ++ks) {
        const float* pp = sm.part[tl * pl.ksplit + ks] + s * SD_PART_LD + half * 8;
        const float4 p0 = *reinterpret_cast<const float4*>(pp), p1 = *reinterpret_cast<const float4*>(pp + 4);
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      emit8(r, s, n, v);
    }
    stp.mark_fine();
    __syncthreads();
  }
}

// y = LN(x (+ add)) * gamma + beta (+ pos) per row (a warp per row, SD_F columns); x, add in shared memory; fp32 result
// into xf, bf16 copy into xs.  x may alias xf.
__device__ void sd_ln_rows(SmemD& sm, const float* x, const float* add, int S, const float* __restrict__ gamma_v,
                           const float* __restrict__ beta_v, const float* __restrict__ gamma_a, const float* __restrict__ beta_a,
                           int T_split, const float* __restrict__ pos) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += SD_WARPS) {
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      float4 a = *reinterpret_cast<const float4*>(x + s * SD_F + c);
      if (add != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(add + s * SD_F + c);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
      sum += a.x + a.y + a.z + a.w;
    }
    const float mean = warp_sum(sum) * (1.f / SD_F);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / SD_F) + SD_EPS);
    const float* gm = s < T_split ? gamma_v : gamma_a;
    const float* bt = s < T_split ? beta_v : beta_a;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(gm + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bt + c));
      float4 o;
      o.x = (v[4 * i] - mean) * rstd * gg.x + bb.x;
      o.y = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
      o.z = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z;
      o.w = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
      if (pos != nullptr) {
        const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + (long long)s * SD_F + c));
        o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
      }
      *reinterpret_cast<float4*>(sm.xf + s * SD_F + c) = o;
      *reinterpret_cast<uint2*>(sm.xs + s * SD_LDX + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

__global__ void __launch_bounds__(SD_THREADS, 1) serve_forward_dsmem_kernel(const ServeParamsD p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemD& sm = *reinterpret_cast<SmemD*>(smem_raw);
  const int rank = (int)sd_rank(), nc = (int)sd_size();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, S = p.S;
  const int64_t* g = p.off_g;
  Stamper stp;
  stp.out = p.stamps;
  stp.on = p.stamps != nullptr && rank == 0 && threadIdx.x == 0;
  stp.fine = false;
  stp.n = 0;
  stp.mark();
  // every CTA of the cluster must be running before the first remote store: arrive now, wait before the first emit
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  const int nF0 = rank * (SD_F / nc), nF1 = nF0 + SD_F / nc;                 // this CTA's slice of a 512-wide output
  const int nQ0 = rank * (3 * SD_F / nc), nQ1 = nQ0 + 3 * SD_F / nc;         // ... of in_proj
  const int nH0 = rank * (p.ffn / nc), nH1 = nH0 + p.ffn / nc;               // ... of linear1
  const int nC0 = rank * (p.hidden / nc), nC1 = nC0 + p.hidden / nc;         // ... of the head's hidden layers
  auto put_xadd = [&](int row0) {
    return [&sm, row0](int r, int s, int n, const float (&v)[8]) { sd_put8_f32(sm.xadd + (row0 + s) * SD_F + n, r, v); };
  };

  // ---- phase 0: input projections (train2.py:150, 153) -> xadd of every CTA.  Video rows in xs, the audio row in
  // row 0 of hs (free until linear1 of layer 0).
  WPre<4> wv, wa;
  sd_prefetch(wv, p.shadow + g[MMER_G_WV], p.video_dim, nF0, nF1);
  sd_prefetch(wa, p.shadow + g[MMER_G_WA], p.audio_dim, nF0, nF1);
  for (int i = threadIdx.x; i < T * (p.video_dim / 8); i += SD_THREADS) {
    const int s = i / (p.video_dim / 8), c = (i % (p.video_dim / 8)) * 8;
    *reinterpret_cast<uint4*>(sm.xs + s * SD_LDX + c) = __ldg(reinterpret_cast<const uint4*>(p.video + (long long)s * p.video_dim + c));
  }
  for (int i = threadIdx.x; i < p.audio_dim / 8; i += SD_THREADS)
    *reinterpret_cast<uint4*>(sm.hs + i * 8) = __ldg(reinterpret_cast<const uint4*>(p.audio + i * 8));
  __syncthreads();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  stp.mark();
  sd_linear(sm, stp, wv, sm.xs, SD_LDX, p.video_dim, p.shadow + g[MMER_G_WV], p.params + g[MMER_G_BV], nF0, nF1, T, false, nc,
            put_xadd(0));
  sd_linear(sm, stp, wa, sm.hs, SD_LDH, p.audio_dim, p.shadow + g[MMER_G_WA], p.params + g[MMER_G_BA], nF0, nF1, 1, false, nc,
            put_xadd(T));
  WPre<8> w;
  sd_prefetch(w, p.shadow + p.off_l[0][MMER_L_IN_W], SD_F, nQ0, nQ1);
  sd_cluster_sync();                                                       // B1
  stp.mark();

  // ---- token assembly (train2.py:151-160) from the local copy
  sd_ln_rows(sm, sm.xadd, nullptr, S, p.params + g[MMER_G_NV_W], p.params + g[MMER_G_NV_B], p.params + g[MMER_G_NA_W],
             p.params + g[MMER_G_NA_B], T, p.params + g[MMER_G_POS]);
  __syncthreads();
  stp.mark();

  const int d = SD_F / p.heads;   // 64
  for (int l = 0; l < p.layers; ++l) {
    const int64_t* o = p.off_l[l];
    stp.fine = p.stamps != nullptr && rank == 0 && threadIdx.x == 0 && l == 0;
    // ---- in_proj: q / k / v rows land per head in every CTA's qkv buffer (last read two barriers ago)
    sd_linear(sm, stp, w, sm.xs, SD_LDX, SD_F, p.shadow + o[MMER_L_IN_W], p.params + o[MMER_L_IN_B], nQ0, nQ1, S, false, nc,
              [&](int r, int s, int n, const float (&v)[8]) {
                const int which = n / SD_F, hc = n - which * SD_F, h = hc >> 6, cc = hc & 63;
                sd_put8_bf16(sm.qkv[h] + which * SD_ROWS * 64 + s * 64 + cc, r, v);
              });
    sd_prefetch(w, p.shadow + o[MMER_L_OUT_W], SD_F, nF0, nF1);
    sd_cluster_sync();                                                     // B2
    stp.mark();
    // ---- attention, one head per warp
    for (int h = warp; h < p.heads; h += SD_WARPS) {
      const bf16* qh = sm.qkv[h];
      const bf16* kh = qh + SD_ROWS * 64;
      const bf16* vh = kh + SD_ROWS * 64;
      float* ps = sm.sc[h];
      for (int idx = lane; idx < S * S; idx += 32) {
        const int i = idx / S, j = idx - i * S;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          const uint4 qa = *reinterpret_cast<const uint4*>(qh + i * 64 + c);
          const uint4 kb = *reinterpret_cast<const uint4*>(kh + j * 64 + c);
          const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qa);
          const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = __bfloat1622float2(q2[e]), b = __bfloat1622float2(k2[e]);
            acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
          }
        }
        const bool masked = (j < T) && p.mask != nullptr && p.mask[j] != 0;
        ps[i * SD_ROWS + j] = masked ? -INFINITY : acc * rsqrtf((float)d);
      }
      __syncwarp();
      if (lane < S) {
        float mx = -INFINITY;
        for (int j = 0; j < S; ++j) mx = fmaxf(mx, ps[lane * SD_ROWS + j]);
        float den = 0.f;
        for (int j = 0; j < S; ++j) { const float e = __expf(ps[lane * SD_ROWS + j] - mx); ps[lane * SD_ROWS + j] = e; den += e; }
        const float inv = 1.f / den;
        for (int j = 0; j < S; ++j) ps[lane * SD_ROWS + j] *= inv;
      }
      __syncwarp();
      for (int i = 0; i < S; ++i) {
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < S; ++j) {
          const float pj = ps[i * SD_ROWS + j];
          const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vh + j * 64 + 2 * lane));
          a0 = fmaf(pj, vv.x, a0);
          a1 = fmaf(pj, vv.y, a1);
        }
        *reinterpret_cast<uint32_t*>(sm.att + i * SD_LDA + h * d + 2 * lane) = pack_bf16x2(a0, a1);
      }
    }
    __syncthreads();
    stp.mark();
    // ---- out_proj -> xadd of every CTA (last read before B2), then x = norm1(x + attention)
    sd_linear(sm, stp, w, sm.att, SD_LDA, SD_F, p.shadow + o[MMER_L_OUT_W], p.params + o[MMER_L_OUT_B], nF0, nF1, S, false, nc,
              put_xadd(0));
    sd_prefetch(w, p.shadow + o[MMER_L_FF1_W], SD_F, nH0, nH1);
    sd_cluster_sync();                                                     // B3
    stp.mark();
    sd_ln_rows(sm, sm.xf, sm.xadd, S, p.params + o[MMER_L_N1_W], p.params + o[MMER_L_N1_B], nullptr, nullptr, S, nullptr);
    __syncthreads();
    stp.mark_fine();
    // ---- linear1 + ReLU -> the bf16 input rows of linear2 in every CTA (last read before B5 of the previous layer)
    sd_linear(sm, stp, w, sm.xs, SD_LDX, SD_F, p.shadow + o[MMER_L_FF1_W], p.params + o[MMER_L_FF1_B], nH0, nH1, S, true, nc,
              [&](int r, int s, int n, const float (&v)[8]) { sd_put8_bf16(sm.hs + s * SD_LDH + n, r, v); });
    sd_prefetch(w, p.shadow + o[MMER_L_FF2_W], p.ffn, nF0, nF1);
    sd_cluster_sync();                                                     // B4
    stp.mark();
    // ---- linear2 -> xadd (last read before B4), then x = norm2(x + ff)
    sd_linear(sm, stp, w, sm.hs, SD_LDH, p.ffn, p.shadow + o[MMER_L_FF2_W], p.params + o[MMER_L_FF2_B], nF0, nF1, S, false, nc,
              put_xadd(0));
    if (l + 1 < p.layers) sd_prefetch(w, p.shadow + p.off_l[l + 1][MMER_L_IN_W], SD_F, nQ0, nQ1);
    else sd_prefetch(w, p.shadow + g[MMER_G_C0_W], SD_F, nC0, nC1);
    sd_cluster_sync();                                                     // B5
    stp.mark();
    sd_ln_rows(sm, sm.xf, sm.xadd, S, p.params + o[MMER_L_N2_W], p.params + o[MMER_L_N2_B], nullptr, nullptr, S, nullptr);
    __syncthreads();
    stp.mark_fine();
  }
  stp.fine = false;

  // ---- masked mean pooling + out_norm (train2.py:184-191): row 0 of xf / xs becomes the fused embedding
  {
    float cnt = 0.f;
    for (int s = 0; s < S; ++s) cnt += ((s < T) && p.mask != nullptr && p.mask[s] != 0) ? 0.f : 1.f;
    const float inv = 1.f / fmaxf(cnt, 1e-6f);
    float* pooled = sm.nrm[0];
    for (int c = threadIdx.x; c < SD_F; c += SD_THREADS) {
      float a = 0.f;
      for (int s = 0; s < S; ++s)
        if (!((s < T) && p.mask != nullptr && p.mask[s] != 0)) a += sm.xf[s * SD_F + c];
      pooled[c] = a * inv;
    }
    __syncthreads();
    sd_ln_rows(sm, pooled, nullptr, 1, p.params + g[MMER_G_ON_W], p.params + g[MMER_G_ON_B], nullptr, nullptr, 1, nullptr);
    __syncthreads();
  }
  stp.mark();
  // ---- classifier head (train2.py:217-229): vectors travel through hbuf (their own buffers: xadd may still be read by a
  // slow CTA's norm2 when a fast one is already here)
  const int Hd = p.hidden;
  auto head_norm = [&](const float* src, const float* gm, const float* bt) {
    float* buf = sm.nrm[0];
    float* g_s = sm.nrm[1];
    float* b_s = sm.nrm[2];
    for (int c = threadIdx.x; c < Hd; c += SD_THREADS) {
      buf[c] = src[c];
      g_s[c] = __ldg(gm + c);
      b_s[c] = __ldg(bt + c);
    }
    __syncthreads();
    if (warp == 0) {
      float sum = 0.f;
      for (int c = lane; c < Hd; c += 32) sum += buf[c];
      const float mean = warp_sum(sum) / (float)Hd;
      float q = 0.f;
      for (int c = lane; c < Hd; c += 32) { const float dd = buf[c] - mean; q = fmaf(dd, dd, q); }
      const float rstd = rsqrtf(warp_sum(q) / (float)Hd + SD_EPS);
      for (int c = lane; c < Hd; c += 32) {
        const float y = fmaxf((buf[c] - mean) * rstd * g_s[c] + b_s[c], 0.f);
        sm.hs[c] = __float2bfloat16_rn(y);       // row 0 of the wide input buffer
        buf[c] = y;
      }
    }
    __syncthreads();
  };
  sd_linear(sm, stp, w, sm.xs, SD_LDX, SD_F, p.shadow + g[MMER_G_C0_W], p.params + g[MMER_G_C0_B], nC0, nC1, 1, false, nc,
            [&](int r, int, int n, const float (&v)[8]) { sd_put8_f32(sm.hbuf[0] + n, r, v); });
  sd_prefetch(w, p.shadow + g[MMER_G_C4_W], Hd, nC0, nC1);
  sd_cluster_sync();
  stp.mark();
  head_norm(sm.hbuf[0], p.params + g[MMER_G_C1_W], p.params + g[MMER_G_C1_B]);
  sd_linear(sm, stp, w, sm.hs, SD_LDH, Hd, p.shadow + g[MMER_G_C4_W], p.params + g[MMER_G_C4_B], nC0, nC1, 1, false, nc,
            [&](int r, int, int n, const float (&v)[8]) { sd_put8_f32(sm.hbuf[1] + n, r, v); });
  sd_cluster_sync();
  stp.mark();
  if (rank == 0) {
    head_norm(sm.hbuf[1], p.params + g[MMER_G_C5_W], p.params + g[MMER_G_C5_B]);
    const float* W8 = p.params + g[MMER_G_C8_W];
    const float* h2 = sm.nrm[0];
    for (int c = warp; c < p.classes; c += SD_WARPS) {
      float a = 0.f;
#pragma unroll 16
      for (int k = lane; k < Hd; k += 32) a = fmaf(h2[k], __ldg(W8 + (long long)c * Hd + k), a);
      a = warp_sum(a);
      if (lane == 0) sm.red[c] = a + __ldg(p.params + g[MMER_G_C8_B] + c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float mx = -INFINITY;
      for (int c = 0; c < p.classes; ++c) mx = fmaxf(mx, sm.red[c]);
      float den = 0.f;
      for (int c = 0; c < p.classes; ++c) den += expf(sm.red[c] - mx);
      for (int c = 0; c < p.classes; ++c) {
        p.logits[c] = sm.red[c];
        p.probs[c] = expf(sm.red[c] - mx) / den;
      }
    }
  }
  stp.mark();
  // no CTA may exit while a peer can still store into its shared memory or wait at a cluster barrier
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// returns 1 when this kernel does not apply (caller falls back to the global-scratch version), 0 on success, < 0 on error
int serve_forward_dsmem(const mmer_model* m, long long* stamps, cudaStream_t st) {
  if (m->T + 1 > SD_ROWS || m->audio_dim > SD_KX || m->video_dim > SD_KX || m->ffn > SD_KH || m->hidden > SD_KH) return 1;
  auto kern = serve_forward_dsmem_kernel;
  static unsigned long long attr_done = 0ull;
  static int cluster = 0;
  const size_t smem = sizeof(SmemD);
  if (needs_func_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(serve dsmem smem)");
    cluster = 8;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16);
      cfg.blockDim = dim3(SD_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) cluster = 16;
    }
    (void)cudaGetLastError();
  }
  ServeParamsD p;
  p.T = m->T; p.S = m->T + 1;
  p.video_dim = m->video_dim; p.audio_dim = m->audio_dim; p.ffn = m->ffn; p.hidden = m->hidden; p.classes = m->classes;
  p.layers = m->layers; p.heads = m->heads;
  p.shadow = reinterpret_cast<const bf16*>(m->shadow);
  p.params = m->params;
  for (int i = 0; i < MMER_G_COUNT; ++i) p.off_g[i] = m->off_g[i];
  for (int l = 0; l < MMER_MAX_LAYERS; ++l)
    for (int i = 0; i < MMER_L_COUNT; ++i) p.off_l[l][i] = m->off_l[l][i];
  p.video = reinterpret_cast<const bf16*>(m->video);
  p.audio = reinterpret_cast<const bf16*>(m->audio);
  p.mask = m->has_mask ? m->mask : nullptr;
  p.stamps = stamps;
  p.logits = m->logits; p.probs = m->probs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cluster);
  cfg.blockDim = dim3(SD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(serve_forward_dsmem)");
  MMER_LAUNCH_CHECK("serve_forward_dsmem_kernel");
  return 0;
}

}  // namespace mmer
