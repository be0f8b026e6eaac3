// Batch-1 serving forward for the live request shape (S = T + 1 <= 8 tokens, train2.py default widths): the same
// one-cluster walk as serve.cu with the work cut so that a transformer layer needs TWO cluster barriers, not four.
//
//   * attention is head-local: the CTA pair (2h, 2h+1) computes q, k, v of head h itself (in_proj rows of that head only;
//     the pair duplicates 96 KB of weight reads, nothing is exchanged), runs the 8x8 attention on tensor cores in one
//     warp, and multiplies its HALF of the head's output columns into out_proj as a split-K partial product over all
//     512 output features;
//   * the feed-forward block is slice-local: a CTA keeps its 128 columns of relu(linear1) in shared memory and multiplies
//     them into linear2 as a split-K partial product;
//   * the 16 partial products of a sub-layer meet in a [512 features][8 tokens] fp32 accumulator in L2 through
//     red.global.add.v2.f32 issued straight from the MMA accumulator fragments (a warp instruction covers two full
//     128-byte lines), then one cluster barrier, then every CTA reads the 16 KB sum and does residual + LayerNorm itself.
//   * the weight fragments of the GEMV that follows a barrier are requested between its arrive and its wait (they do
//     not depend on activations): the loads fly while the slowest CTA arrives;
//   * every bias / LayerNorm vector is staged in shared memory at start, so no phase begins with an L2 round trip for
//     parameters;
//   * the weights are read from a copy packed in MMA-fragment order (mmer_serve_pack): the 16 rows x 32 columns a warp
//     needs for one k-step are 1 KB contiguous, so each 16-byte-per-lane load covers four whole 128-byte lines instead of
//     eight half lines (the loads, not the MMAs, pace every phase: ~1 us per 128 KB per SM in the row-major layout).
//
// The fp32 adds arrive in arbitrary order, so two identical calls may differ in the last bits (bf16 flips downstream);
// serve.cu (MMER_DEBUG_SERVE_GLOBAL) is the run-to-run bit-reproducible version.
// Input projections and the classifier head keep the output-feature split with an L2 exchange (their LayerNorms need
// whole rows).  Barriers: 1 + 2 per layer + 2.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

extern int g_debug[16];

namespace {

constexpr int SS_THREADS = 512;
constexpr int SS_WARPS = 16;
constexpr int SS_NC = 16;                // CTAs per cluster (fixed: the cut above is written for 2 CTAs per head)
constexpr int SS_F = 512;
constexpr int SS_HEADS = 8;
constexpr int SS_D = 64;
constexpr int SS_FFN = 2048;
constexpr int SS_HS = SS_FFN / SS_NC;    // 128 hidden columns per CTA
constexpr int SS_KX = 1024;              // widest row held in xs
constexpr int SS_LDX = SS_KX + 32;       // 64-byte skew: conflict-free 16-byte B-fragment loads
constexpr int SS_LDH = SS_HS + 32;
constexpr int SS_LDO = SS_D + 32;
constexpr int SS_LDQ = SS_D + 8;         // q / k rows: conflict-free 4-byte fragment loads
constexpr int SS_LDS = SS_F + 4;         // fp32 sub-layer sum rows
constexpr int SS_PART_LD = 20;
constexpr int SS_MAXH = 2048;            // classifier hidden width
constexpr int SS_MAXL = 4;              // layers (the kernel parameter block stays small: it is read cold at every launch)
constexpr int SS_LN_STAGED = 2;         // layers whose LayerNorm / bias vectors are staged in shared memory at kernel start
constexpr float SS_EPS = 1e-5f;
constexpr int SS_HN = 512;               // head LayerNorm vectors are staged up to this hidden width

// sm.sv (floats): every small vector a phase would otherwise fetch from L2 on its critical path
constexpr int SV_BV = 0;                          // [32] video projection bias slice
constexpr int SV_BA = SV_BV + 32;                 // [32] audio projection bias slice
constexpr int SV_C0B = SV_BA + 32;                // [<= 128] head linear 0 bias slice
constexpr int SV_C4B = SV_C0B + 128;              // [<= 128] head linear 1 bias slice
constexpr int SV_FF1B = SV_C4B + 128;             // [SS_MAXL][128] linear1 bias slice
constexpr int SV_INB = SV_FF1B + SS_MAXL * 128;   // [SS_MAXL][3][64] in_proj bias of this CTA's head
constexpr int SV_ON = SV_INB + SS_MAXL * 192;     // [2][512] out_norm weight, bias
constexpr int SV_HN = SV_ON + 2 * SS_F;           // [4][SS_HN] head norm 0 weight, bias, head norm 1 weight, bias
constexpr int SV_TOTAL = SV_HN + 4 * SS_HN;

// scratch (floats), inside the buffer of mmer_serve_scratch_bytes()
constexpr int SF_PRE = 0;                // [8][512] projections before LayerNorm
constexpr int SF_H1 = SF_PRE + 8 * SS_F; // [2048]
constexpr int SF_H2 = SF_H1 + SS_MAXH;
constexpr int SF_ACC = SF_H2 + SS_MAXH;  // [2 * layers][512][8] split-K sums
constexpr int SF_ACC_SZ = SS_F * 8;

struct ServeParamsS {
  int T, S;
  int video_dim, audio_dim, hidden, classes, layers;
  const bf16* shadow;                  // the PACKED copy (mmer_serve_pack), same offsets as the row-major shadow
  const float* params;
  int64_t off_g[MMER_G_COUNT];
  int64_t off_l[SS_MAXL][MMER_L_COUNT];
  const bf16* video;
  const bf16* audio;
  const uint8_t* mask;
  float* scratch;
  long long* stamps;
  int fine;                            // MMER_DEBUG_SERVE_STAMPS: extra marks inside phase 0 and layer 0
  float* logits;
  float* probs;
};

struct SmemS {
  bf16 xs[8 * SS_LDX];                 // GEMV input rows (video rows, then the bf16 residual stream)
  bf16 arow[SS_KX + SS_MAXH];          // single-row GEMV inputs: audio features; head vectors
  float xf[8 * SS_F];                  // residual stream, fp32
  float sum[8 * SS_LDS];               // the sub-layer output read back from the L2 accumulator, [token][feature]
  bf16 q[8 * SS_LDQ];                  // this CTA's head: q, k as [token][feature]
  bf16 k[8 * SS_LDQ];
  bf16 vt[SS_D * 8];                   // v as [feature][token]
  bf16 ao[8 * SS_LDO];                 // attention output of the head, [token][feature]
  bf16 hs[8 * SS_LDH];                 // relu(linear1) slice, [token][128]
  float part[SS_WARPS][8 * SS_PART_LD];
  uint4 wpark[12 * 16 * 32];           // in_proj weight fragments of k-steps 8..15, parked by cp.async ahead of the barrier
  float lnp[SS_LN_STAGED][6][SS_F];     // out_proj bias, norm1 w / b, linear2 bias, norm2 w / b of the first layers
  float sv[SV_TOTAL];                  // this CTA's slices of every other bias / norm vector (offsets SV_*), staged at start
  float nrm[SS_MAXH];                  // pooled embedding; head hidden vector after its LayerNorm
  float red[64];
};

__device__ __forceinline__ uint32_t ss_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// The cluster barrier in its two halves.  arrive: this thread's global writes and reductions are released to the
// cluster.  The caller then issues the NEXT phase's weight loads (they depend on no activation), then waits: the loads
// fly while the slowest CTA arrives.  (Issued before the arrive they were waited for by its release fence: 1.1 us per
// barrier.)  Every thread arrives itself, so the barrier also orders this CTA's shared memory: no __syncthreads.
__device__ __forceinline__ void ss_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void ss_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// a weight load that stays where it is written (ahead of the barrier), not where its value is first used
__device__ __forceinline__ uint4 ss_ldw(const bf16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ss_red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// One accumulator fragment (feature g: tokens 2t, 2t+1 in c0, c1; feature g+8 in c2, c3) into the L2 sum
// tile[16 features][8 tokens]: neighbouring lanes swap halves so that each issues ONE 16-byte reduction (half the
// instructions of the 8-byte form; measured equal in time): even t -> feature g, tokens 2t..2t+3; odd t -> feature g+8, tokens 2t-2..2t+1.
__device__ __forceinline__ void ss_red_tile(float* tile, const float (&c)[4], int S) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool odd = t & 1;
  const float r0 = __shfl_xor_sync(0xffffffffu, odd ? c[0] : c[2], 1);
  const float r1 = __shfl_xor_sync(0xffffffffu, odd ? c[1] : c[3], 1);
  const int tok0 = odd ? 2 * t - 2 : 2 * t;
  if (tok0 < S) {
    if (!odd) ss_red4(tile + g * 8 + tok0, c[0], c[1], r0, r1);
    else ss_red4(tile + (g + 8) * 8 + tok0, r0, r1, c[2], c[3]);
  }
}
// D(16x8) += A(16x8, bf16, row) . B(8x8, bf16, col)
__device__ __forceinline__ void ss_mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

struct StampS {
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

struct WPre8 {
  uint4 a0[8], a1[8];
};

// Packed weights: block (tile, k-step) = rows [16 tile, 16 tile + 16) x columns [32 ks, 32 ks + 32) of W[N][K], 512
// elements: lane L = 4 g + t holds row g, columns t*8..t*8+7 at L*8 (fragment a0) and row g + 8 at 256 + L*8 (a1).
__device__ __forceinline__ const bf16* ss_wp(const bf16* Wp, int K, int tile, int ks) {
  return Wp + ((long long)(tile * (K >> 5) + ks) << 9) + (threadIdx.x & 31) * 8;
}

// ---------------------------------------------------------------------------------------------------------------------
// (1) output-feature split with a k-split across warps: a warp owns (tile, k-range); partial tiles meet in shared memory.
struct PlanS {
  int tiles, ksplit, klen, units;
};
__device__ __forceinline__ PlanS ss_plan(int K, int n0, int n1) {
  PlanS pl;
  pl.tiles = (n1 - n0) >> 4;
  pl.ksplit = 1;
  while (pl.ksplit * 2 * pl.tiles <= SS_WARPS && (K / (pl.ksplit * 2)) % 32 == 0) pl.ksplit *= 2;
  pl.units = pl.tiles * pl.ksplit;
  pl.klen = K / pl.ksplit;
  return pl;
}
// The first NSLOTS k-steps of this warp's unit go into prefetch slots [SLOT0, SLOT0 + NSLOTS): two GEMVs with short
// k-ranges (the input projections) can share one register set.
template <int SLOT0 = 0, int NSLOTS = 8>
__device__ __forceinline__ void ss_nsplit_prefetch(WPre8& w, const bf16* __restrict__ W, int K, int n0, int n1) {
  static_assert(SLOT0 + NSLOTS <= 8, "prefetch slots");
  const int warp = threadIdx.x >> 5;
  const PlanS pl = ss_plan(K, n0, n1);
  if (warp < pl.units) {
    const int tile = warp / pl.ksplit, ks = warp - tile * pl.ksplit;
    const bf16* w0 = ss_wp(W, K, (n0 >> 4) + tile, (ks * pl.klen) >> 5);
#pragma unroll
    for (int i = 0; i < NSLOTS; ++i)
      if (i * 32 < pl.klen) {
        w.a0[SLOT0 + i] = ss_ldw(w0 + i * 512);
        w.a1[SLOT0 + i] = ss_ldw(w0 + i * 512 + 256);
      }
  }
}
// Part 1: this warp's (tile, k-range) partial tile -> sm.part.  x rows of ldx elements in shared memory (ldx = 0: one
// row for every token slot).  One round: tiles * ksplit <= 16 (output slices of at most 256 features).
template <int SLOT0 = 0, int NSLOTS = 8>
__device__ __forceinline__ void ss_nsplit_mma(SmemS& sm, const WPre8& pre, const bf16* xs, int ldx, int K, const bf16* __restrict__ W,
                                              int n0, int n1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const PlanS pl = ss_plan(K, n0, n1);
  if (warp < pl.units) {
    const int tile = warp / pl.ksplit, ks = warp - tile * pl.ksplit;
    const bf16* w0 = ss_wp(W, K, (n0 >> 4) + tile, (ks * pl.klen) >> 5);
    const bf16* x0 = xs + g * ldx + ks * pl.klen + t * 8;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NSLOTS; ++i)
      if (i * 32 < pl.klen) {
        const uint4 b = *reinterpret_cast<const uint4*>(x0 + i * 32);
        const uint4 a0 = pre.a0[SLOT0 + i], a1 = pre.a1[SLOT0 + i];
        mma_bf16_16816(c, a0.x, a1.x, a0.y, a1.y, b.x, b.y);
        mma_bf16_16816(c, a0.z, a1.z, a0.w, a1.w, b.z, b.w);
      }
#pragma unroll 4
    for (int kb = NSLOTS * 32; kb < pl.klen; kb += 32) {  // k-steps beyond the prefetched ones
      const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(w0 + (kb >> 5) * 512));
      const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(w0 + (kb >> 5) * 512 + 256));
      const uint4 b = *reinterpret_cast<const uint4*>(x0 + kb);
      mma_bf16_16816(c, a0.x, a1.x, a0.y, a1.y, b.x, b.y);
      mma_bf16_16816(c, a0.z, a1.z, a0.w, a1.w, b.z, b.w);
    }
    float* pp = sm.part[warp];       // [token][feature]
    pp[(2 * t) * SS_PART_LD + g] = c[0];
    pp[(2 * t + 1) * SS_PART_LD + g] = c[1];
    pp[(2 * t) * SS_PART_LD + g + 8] = c[2];
    pp[(2 * t + 1) * SS_PART_LD + g + 8] = c[3];
  }
}
// Part 2: emit8(s, n, v[8]) for token s < S and each 8-feature group n of [n0, n1): v = act(x[s] . W[n..n+7] + bias)
// bias: this CTA's slice (element 0 = feature n0), shared memory.
template <typename Emit>
__device__ __forceinline__ void ss_nsplit_reduce(SmemS& sm, int K, const float* bias, int n0, int n1, int S, bool relu,
                                                 Emit emit8) {
  const PlanS pl = ss_plan(K, n0, n1);
  __syncthreads();
  const int groups = pl.tiles * 2;
  for (int e = threadIdx.x; e < groups * S; e += SS_THREADS) {
    const int grp = e % groups, s = e / groups;
    const int tl = grp >> 1, half = grp & 1;
    const int n = n0 + tl * 16 + half * 8;
    float v[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias + (n - n0)), b1 = *reinterpret_cast<const float4*>(bias + (n - n0) + 4);
    v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    for (int ks = 0; ks < pl.ksplit; ++ks) {
      const float* pp = sm.part[tl * pl.ksplit + ks] + s * SS_PART_LD + half * 8;
      const float4 p0 = *reinterpret_cast<const float4*>(pp), p1 = *reinterpret_cast<const float4*>(pp + 4);
      v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    emit8(s, n, v);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
// (2) split-K partial product over all 512 output features: a warp owns two 16-feature tiles; the k-slice [k0, k0 + 32 ITERS)
// of W's columns against this CTA's activation slice (rows of ldx elements, first element = column k0).  Accumulator
// fragments go to the L2 sum [feature][8 tokens] as 16-byte vector reductions.
template <int ITERS>
__device__ __forceinline__ void ss_ksplit_prefetch(WPre8& w, const bf16* __restrict__ W, int K, int k0) {
  static_assert(2 * ITERS <= 8, "prefetch slots");
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    const bf16* w0 = ss_wp(W, K, warp * 2 + tl, k0 >> 5);
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      w.a0[tl * ITERS + i] = ss_ldw(w0 + i * 512);
      w.a1[tl * ITERS + i] = ss_ldw(w0 + i * 512 + 256);
    }
  }
}
template <int ITERS>
__device__ __forceinline__ void ss_ksplit(const WPre8& w, const bf16* xs, int ldx, float* __restrict__ acc, int S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint4 b[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) b[i] = *reinterpret_cast<const uint4*>(xs + g * ldx + i * 32 + t * 8);
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const uint4 a0 = w.a0[tl * ITERS + i], a1 = w.a1[tl * ITERS + i];
      mma_bf16_16816(c, a0.x, a1.x, a0.y, a1.y, b[i].x, b[i].y);
      mma_bf16_16816(c, a0.z, a1.z, a0.w, a1.w, b[i].z, b[i].w);
    }
    ss_red_tile(acc + (warp * 2 + tl) * 128, c, S);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// (3) in_proj rows of one head: 12 tiles (q, k, v x 64 features), warp i < 12 owns tile i over the whole K = 512;
// the first 8 k-steps come from the prefetch.
__device__ __forceinline__ const bf16* ss_head_tile(const bf16* Wp, int h, int tile) {
  const int which = tile >> 2, sub = tile & 3;
  return ss_wp(Wp, SS_F, (which * SS_F + h * SS_D + sub * 16) >> 4, 0);
}
__device__ __forceinline__ void ss_head_prefetch(SmemS& sm, WPre8& w, const bf16* __restrict__ W, int h) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 12) {
    const bf16* w0 = ss_head_tile(W, h, warp);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w.a0[i] = ss_ldw(w0 + i * 512);
      w.a1[i] = ss_ldw(w0 + i * 512 + 256);
    }
    // k-steps 8..15: each lane parks its own future fragments in shared memory (no register room for them)
    uint4* park = sm.wpark + (warp * 16) * 32 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cp_async_16(smem_u32(park + (2 * i) * 32), w0 + (8 + i) * 512);
      cp_async_16(smem_u32(park + (2 * i + 1) * 32), w0 + (8 + i) * 512 + 256);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");   // every thread: group counts stay uniform across the CTA
}
__device__ __forceinline__ void ss_head_qkv(SmemS& sm, const WPre8& w, const float* bias /* staged [3][64] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (warp >= 12) return;
  const bf16* x0 = sm.xs + g * SS_LDX + t * 8;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 b = *reinterpret_cast<const uint4*>(x0 + i * 32);
    mma_bf16_16816(c, w.a0[i].x, w.a1[i].x, w.a0[i].y, w.a1[i].y, b.x, b.y);
    mma_bf16_16816(c, w.a0[i].z, w.a1[i].z, w.a0[i].w, w.a1[i].w, b.z, b.w);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const uint4* park = sm.wpark + (warp * 16) * 32 + lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a0 = park[(2 * i) * 32], a1 = park[(2 * i + 1) * 32];
    const uint4 b = *reinterpret_cast<const uint4*>(x0 + (8 + i) * 32);
    mma_bf16_16816(c, a0.x, a1.x, a0.y, a1.y, b.x, b.y);
    mma_bf16_16816(c, a0.z, a1.z, a0.w, a1.w, b.z, b.w);
  }
  const int which = warp >> 2, f = (warp & 3) * 16 + g;            // feature of the head (and f + 8)
  const float b_lo = bias[which * SS_D + f], b_hi = bias[which * SS_D + f + 8];
  c[0] += b_lo; c[1] += b_lo; c[2] += b_hi; c[3] += b_hi;          // (f, token 2t), (f, 2t+1), (f+8, 2t), (f+8, 2t+1)
  if (which == 2) {
    *reinterpret_cast<uint32_t*>(sm.vt + f * 8 + 2 * t) = pack_bf16x2(c[0], c[1]);
    *reinterpret_cast<uint32_t*>(sm.vt + (f + 8) * 8 + 2 * t) = pack_bf16x2(c[2], c[3]);
  } else {
    bf16* dst = which == 0 ? sm.q : sm.k;
    dst[(2 * t) * SS_LDQ + f] = __float2bfloat16_rn(c[0]);
    dst[(2 * t + 1) * SS_LDQ + f] = __float2bfloat16_rn(c[1]);
    dst[(2 * t) * SS_LDQ + f + 8] = __float2bfloat16_rn(c[2]);
    dst[(2 * t + 1) * SS_LDQ + f + 8] = __float2bfloat16_rn(c[3]);
  }
}

// (4) softmax(q k^T / 8 + key mask) v for one head on tensor cores, one warp.  Token slots >= S hold finite values
// (rows of xs beyond S are zero) and are switched off by select, never by arithmetic.
__device__ __forceinline__ void ss_attention(SmemS& sm, unsigned valid) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k0 = 0; k0 < SS_D; k0 += 16) {
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(sm.q + g * SS_LDQ + k0 + 2 * t);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(sm.q + g * SS_LDQ + k0 + 8 + 2 * t);
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sm.k + g * SS_LDQ + k0 + 2 * t);
    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sm.k + g * SS_LDQ + k0 + 8 + 2 * t);
    mma_bf16_16816(c, a0, 0u, a2, 0u, b0, b1);
  }
  // c[0], c[1]: query g, keys 2t, 2t+1
  const int j0 = 2 * t, j1 = 2 * t + 1;
  const bool off0 = !((valid >> j0) & 1u), off1 = !((valid >> j1) & 1u);
  const float s0 = off0 ? -INFINITY : c[0] * 0.125f, s1 = off1 ? -INFINITY : c[1] * 0.125f;
  float mx = fmaxf(s0, s1);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  const float e0 = off0 ? 0.f : __expf(s0 - mx), e1 = off1 ? 0.f : __expf(s1 - mx);
  float den = e0 + e1;
  den += __shfl_xor_sync(0xffffffffu, den, 1);
  den += __shfl_xor_sync(0xffffffffu, den, 2);
  const float inv = 1.f / den;
  const uint32_t pa = pack_bf16x2(e0 * inv, e1 * inv);             // A fragment of m16n8k8: (row g, k 2t, 2t+1)
#pragma unroll
  for (int n0 = 0; n0 < SS_D; n0 += 8) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sm.vt + (n0 + g) * 8 + 2 * t);   // (k 2t, 2t+1; column n0+g)
    ss_mma_1688(o, pa, 0u, b0);
    *reinterpret_cast<uint32_t*>(sm.ao + g * SS_LDO + n0 + 2 * t) = pack_bf16x2(o[0], o[1]);
  }
}

// y = LN(x + add + bias) * gamma + beta (+ pos) per row (a warp per row); fp32 -> xf, bf16 -> xs.  x / add in shared
// memory or (x_global) the L2 scratch.
template <bool PRELOAD>
__device__ void ss_ln_rows(SmemS& sm, const float* x, int ldx, bool x_global, const float* add, const float* bias, int S,
                           const float* gamma_v, const float* beta_v, const float* gamma_a, const float* beta_a, int T_split,
                           const float* __restrict__ pos) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += SS_WARPS) {
    const float* gm = s < T_split ? gamma_v : gamma_a;
    const float* bt = s < T_split ? beta_v : beta_a;
    float4 gg[4], bb[4];
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // PRELOAD (vectors in global memory): every load before the first reduction
      const int c = i * 128 + lane * 4;
      if (PRELOAD) {
        gg[i] = *reinterpret_cast<const float4*>(gm + c);
        bb[i] = *reinterpret_cast<const float4*>(bt + c);
        if (pos != nullptr) {
          const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + (long long)s * SS_F + c));
          bb[i].x += pp.x; bb[i].y += pp.y; bb[i].z += pp.z; bb[i].w += pp.w;
        }
      }
      float4 a = x_global ? __ldcg(reinterpret_cast<const float4*>(x + s * ldx + c)) : *reinterpret_cast<const float4*>(x + s * ldx + c);
      if (add != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(add + s * SS_LDS + c);
        const float4 ob = *reinterpret_cast<const float4*>(bias + c);
        a.x += b.x + ob.x; a.y += b.y + ob.y; a.z += b.z + ob.z; a.w += b.w + ob.w;
      }
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += v[i];
    const float mean = warp_sum(sum) * (1.f / SS_F);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / SS_F) + SS_EPS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      if (!PRELOAD) {
        gg[i] = *reinterpret_cast<const float4*>(gm + c);
        bb[i] = *reinterpret_cast<const float4*>(bt + c);
      }
      float4 o;
      o.x = (v[4 * i] - mean) * rstd * gg[i].x + bb[i].x;
      o.y = (v[4 * i + 1] - mean) * rstd * gg[i].y + bb[i].y;
      o.z = (v[4 * i + 2] - mean) * rstd * gg[i].z + bb[i].z;
      o.w = (v[4 * i + 3] - mean) * rstd * gg[i].w + bb[i].w;
      *reinterpret_cast<float4*>(sm.xf + s * SS_F + c) = o;
      *reinterpret_cast<uint2*>(sm.xs + s * SS_LDX + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

// the L2 sum [feature][8 tokens] -> sm.sum [token][feature]; 1024 float4, two per thread, all in flight
__device__ __forceinline__ void ss_fetch_sum(SmemS& sm, const float* __restrict__ acc) {
  float4 v[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(acc) + threadIdx.x + u * SS_THREADS);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = threadIdx.x + u * SS_THREADS, n = i >> 1, s0 = (i & 1) * 4;
    sm.sum[(s0 + 0) * SS_LDS + n] = v[u].x;
    sm.sum[(s0 + 1) * SS_LDS + n] = v[u].y;
    sm.sum[(s0 + 2) * SS_LDS + n] = v[u].z;
    sm.sum[(s0 + 3) * SS_LDS + n] = v[u].w;
  }
}

__global__ void __launch_bounds__(SS_THREADS, 1) serve_small_kernel(const ServeParamsS p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemS& sm = *reinterpret_cast<SmemS*>(smem_raw);
  const int rank = (int)ss_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, S = p.S;
  const int64_t* g = p.off_g;
  float* sc = p.scratch;
  StampS stp;
  stp.out = p.stamps;
  stp.on = p.stamps != nullptr && rank == 0 && threadIdx.x == 0;
  stp.fine = stp.on && (p.fine & 1);        // extra marks through phase 0 and layer 0
  stp.n = 0;
  stp.mark();
  const int nF0 = rank * (SS_F / SS_NC), nF1 = nF0 + SS_F / SS_NC;
  const int nH0 = rank * SS_HS, nH1 = nH0 + SS_HS;
  const int nC0 = rank * (p.hidden / SS_NC), nC1 = nC0 + p.hidden / SS_NC;
  const int head = rank >> 1, khalf = (rank & 1) * 32;

  // ---- phase 0: input projections (train2.py:150, 153), output-feature split; the split-K sums are zeroed here, one
  // share per CTA (the first barrier publishes them)
  WPre8 w;
  const bf16* Wv = p.shadow + g[MMER_G_WV];
  const bf16* Wa = p.shadow + g[MMER_G_WA];
  ss_nsplit_prefetch<0, 4>(w, Wv, p.video_dim, nF0, nF1);     // both input projections' weights now: k-ranges of <= 128 each
  ss_nsplit_prefetch<4, 4>(w, Wa, p.audio_dim, nF0, nF1);
  {
    // the projection biases (needed in this phase) as their own cp.async group
    for (int i = threadIdx.x * 4; i < nF1 - nF0; i += SS_THREADS * 4) {
      cp_async_16(smem_u32(sm.sv + SV_BV + i), p.params + g[MMER_G_BV] + nF0 + i);
      cp_async_16(smem_u32(sm.sv + SV_BA + i), p.params + g[MMER_G_BA] + nF0 + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // key padding mask as bits (bit s set = token s takes part); the audio token always does
  unsigned valid = 0u;
  for (int s = 0; s < S; ++s) valid |= ((s < T && p.mask != nullptr && p.mask[s] != 0) ? 0u : 1u) << s;
  for (int i = threadIdx.x; i < 8 * (p.video_dim / 8); i += SS_THREADS) {
    const int s = i / (p.video_dim / 8), c = (i % (p.video_dim / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (s < T) v = __ldg(reinterpret_cast<const uint4*>(p.video + (long long)s * p.video_dim + c));
    *reinterpret_cast<uint4*>(sm.xs + s * SS_LDX + c) = v;
  }
  for (int i = threadIdx.x; i < p.audio_dim / 8; i += SS_THREADS)
    *reinterpret_cast<uint4*>(sm.arow + i * 8) = __ldg(reinterpret_cast<const uint4*>(p.audio + i * 8));
  {
    const int share = 2 * p.layers * SF_ACC_SZ / 4 / SS_NC;       // float4 per CTA
    float4* z = reinterpret_cast<float4*>(sc + SF_ACC) + (long long)rank * share;
    for (int i = threadIdx.x; i < share; i += SS_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // the projection biases
  __syncthreads();
  stp.mark_fine();
  auto put_rows = [sc](int row0) {
    return [sc, row0](int s, int n, const float (&v)[8]) {
      float4* d = reinterpret_cast<float4*>(sc + SF_PRE + (row0 + s) * SS_F + n);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    };
  };
  ss_nsplit_mma<0, 4>(sm, w, sm.xs, SS_LDX, p.video_dim, Wv, nF0, nF1);
  stp.mark_fine();
  ss_nsplit_reduce(sm, p.video_dim, sm.sv + SV_BV, nF0, nF1, T, false, put_rows(0));
  stp.mark_fine();
  ss_nsplit_mma<4, 4>(sm, w, sm.arow, 0, p.audio_dim, Wa, nF0, nF1);
  ss_nsplit_reduce(sm, p.audio_dim, sm.sv + SV_BA, nF0, nF1, 1, false, put_rows(T));
  // rows >= T of xs held zeros through the video GEMV; from here on rows >= S stay zero (finite q / k / v in the
  // unused token slots): the LayerNorms write rows < S only
  stp.mark_fine();
  ss_cluster_arrive();                                                     // B1
  stp.mark_fine();
  ss_head_prefetch(sm, w, p.shadow + p.off_l[0][MMER_L_IN_W], head);
  {
    // every bias / LayerNorm vector this CTA will need -> shared memory, requested while this CTA waits at the
    // first barrier (instead of an L2 round trip at the head of each later phase)
    static constexpr int which[6] = {MMER_L_OUT_B, MMER_L_N1_W, MMER_L_N1_B, MMER_L_FF2_B, MMER_L_N2_W, MMER_L_N2_B};
    const int nl = min(p.layers, SS_LN_STAGED);
    for (int i = threadIdx.x; i < nl * 6 * (SS_F / 4); i += SS_THREADS) {
      const int l = i / (6 * (SS_F / 4)), r = i - l * 6 * (SS_F / 4), v = r / (SS_F / 4), c = (r - v * (SS_F / 4)) * 4;
      cp_async_16(smem_u32(&sm.lnp[l][v][c]), p.params + p.off_l[l][which[v]] + c);
    }
    auto stage = [&](int dst, const float* src, int count) {       // count % 4 == 0, src 16-byte aligned
      for (int i = threadIdx.x * 4; i < count; i += SS_THREADS * 4) cp_async_16(smem_u32(sm.sv + dst + i), src + i);
    };
    stage(SV_C0B, p.params + g[MMER_G_C0_B] + nC0, nC1 - nC0);
    stage(SV_C4B, p.params + g[MMER_G_C4_B] + nC0, nC1 - nC0);
    for (int l = 0; l < p.layers; ++l) {
      stage(SV_FF1B + l * 128, p.params + p.off_l[l][MMER_L_FF1_B] + nH0, SS_HS);
      for (int q = 0; q < 3; ++q)
        stage(SV_INB + l * 192 + q * SS_D, p.params + p.off_l[l][MMER_L_IN_B] + q * SS_F + head * SS_D, SS_D);
    }
    stage(SV_ON, p.params + g[MMER_G_ON_W], SS_F);
    stage(SV_ON + SS_F, p.params + g[MMER_G_ON_B], SS_F);
    if (p.hidden <= SS_HN) {
      stage(SV_HN, p.params + g[MMER_G_C1_W], p.hidden);
      stage(SV_HN + SS_HN, p.params + g[MMER_G_C1_B], p.hidden);
      stage(SV_HN + 2 * SS_HN, p.params + g[MMER_G_C5_W], p.hidden);
      stage(SV_HN + 3 * SS_HN, p.params + g[MMER_G_C5_B], p.hidden);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  stp.mark_fine();
  ss_cluster_wait();
  stp.mark();

  // ---- token assembly (train2.py:151-160)
  ss_ln_rows<true>(sm, sc + SF_PRE, SS_F, true, nullptr, nullptr, S, p.params + g[MMER_G_NV_W], p.params + g[MMER_G_NV_B],
             p.params + g[MMER_G_NA_W], p.params + g[MMER_G_NA_B], T, p.params + g[MMER_G_POS]);
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // staged vectors and parked in_proj fragments
  __syncthreads();
  stp.mark();

  for (int l = 0; l < p.layers; ++l) {
    const int64_t* o = p.off_l[l];
    float* acc_a = sc + SF_ACC + (2 * l) * SF_ACC_SZ;
    float* acc_f = acc_a + SF_ACC_SZ;
    stp.fine = stp.on && (p.fine & 1) && l == 0;
    const bool staged = l < SS_LN_STAGED;
    const float* b_out = staged ? sm.lnp[l][0] : p.params + o[MMER_L_OUT_B];
    const float* n1w = staged ? sm.lnp[l][1] : p.params + o[MMER_L_N1_W];
    const float* n1b = staged ? sm.lnp[l][2] : p.params + o[MMER_L_N1_B];
    const float* b_ff2 = staged ? sm.lnp[l][3] : p.params + o[MMER_L_FF2_B];
    const float* n2w = staged ? sm.lnp[l][4] : p.params + o[MMER_L_N2_W];
    const float* n2b = staged ? sm.lnp[l][5] : p.params + o[MMER_L_N2_B];
    // ---- q, k, v of this CTA's head; attention; its half of the head's columns times out_proj -> L2 sum
    ss_head_qkv(sm, w, sm.sv + SV_INB + l * 192);
    ss_ksplit_prefetch<1>(w, p.shadow + o[MMER_L_OUT_W], SS_F, head * SS_D + khalf);
    stp.mark_fine();
    __syncthreads();
    if (warp == 0) ss_attention(sm, valid);
    __syncthreads();
    stp.mark_fine();
    ss_ksplit<1>(w, sm.ao + khalf, SS_LDO, acc_a, S);
    stp.mark_fine();
    ss_cluster_arrive();                                                   // B2
    stp.mark_fine();
    ss_nsplit_prefetch(w, p.shadow + o[MMER_L_FF1_W], SS_F, nH0, nH1);
    stp.mark_fine();
    ss_cluster_wait();
    stp.mark();
    // ---- x = norm1(x + out_proj(attention)); relu(linear1) slice stays here; times its linear2 columns -> L2 sum
    ss_fetch_sum(sm, acc_a);
    __syncthreads();
    ss_ln_rows<false>(sm, sm.xf, SS_F, false, sm.sum, b_out, S, n1w, n1b, nullptr, nullptr, S, nullptr);
    __syncthreads();
    stp.mark_fine();
    ss_nsplit_mma(sm, w, sm.xs, SS_LDX, SS_F, p.shadow + o[MMER_L_FF1_W], nH0, nH1);
    ss_ksplit_prefetch<4>(w, p.shadow + o[MMER_L_FF2_W], SS_FFN, nH0);
    stp.mark_fine();
    ss_nsplit_reduce(sm, SS_F, sm.sv + SV_FF1B + l * 128, nH0, nH1, 8, true, [&](int s, int n, const float (&v)[8]) {
      *reinterpret_cast<uint4*>(sm.hs + s * SS_LDH + (n - nH0)) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    });
    stp.mark_fine();
    ss_ksplit<4>(w, sm.hs, SS_LDH, acc_f, S);
    stp.mark_fine();
    ss_cluster_arrive();                                                   // B3
    stp.mark_fine();
    if (l + 1 < p.layers) ss_head_prefetch(sm, w, p.shadow + p.off_l[l + 1][MMER_L_IN_W], head);
    else ss_nsplit_prefetch(w, p.shadow + g[MMER_G_C0_W], SS_F, nC0, nC1);
    stp.mark_fine();
    ss_cluster_wait();
    stp.mark();
    ss_fetch_sum(sm, acc_f);
    __syncthreads();
    ss_ln_rows<false>(sm, sm.xf, SS_F, false, sm.sum, b_ff2, S, n2w, n2b, nullptr, nullptr, S, nullptr);
    __syncthreads();
    stp.mark_fine();
  }
  stp.fine = false;

  // ---- masked mean pooling + out_norm (train2.py:184-191) -> row 0 of xf / xs
  {
    const float inv = 1.f / fmaxf((float)__popc(valid), 1e-6f);
    float* pooled = sm.nrm;
    for (int c = threadIdx.x; c < SS_F; c += SS_THREADS) {
      float a = 0.f;
      for (int s = 0; s < S; ++s)
        if ((valid >> s) & 1u) a += sm.xf[s * SS_F + c];
      pooled[c] = a * inv;
    }
    __syncthreads();
    ss_ln_rows<false>(sm, pooled, SS_F, false, nullptr, nullptr, 1, sm.sv + SV_ON, sm.sv + SV_ON + SS_F, nullptr, nullptr, 1, nullptr);
    __syncthreads();
  }
  stp.mark();
  // ---- classifier head (train2.py:217-229), output-feature split, vectors through the L2 scratch
  const int Hd = p.hidden;
  // LayerNorm + ReLU of a hidden vector (whole CTA, up to 4 elements per thread): bf16 -> arow, fp32 -> nrm
  auto head_norm = [&](const float* src, const float* gm, const float* bt) {
    float v[4], gv[4], bv[4];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = threadIdx.x + j * SS_THREADS;
      v[j] = 0.f;
      if (c < Hd) {
        v[j] = __ldcg(src + c);
        gv[j] = gm[c];
        bv[j] = bt[c];
      }
      sum += v[j];
    }
    sum = warp_sum(sum);
    if (lane == 0) sm.red[warp] = sum;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < SS_WARPS; ++i) mean += sm.red[i];
    mean /= (float)Hd;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (threadIdx.x + j * SS_THREADS < Hd) { const float dd = v[j] - mean; q = fmaf(dd, dd, q); }
    q = warp_sum(q);
    if (lane == 0) sm.red[SS_WARPS + warp] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < SS_WARPS; ++i) var += sm.red[SS_WARPS + i];
    const float rstd = rsqrtf(var / (float)Hd + SS_EPS);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = threadIdx.x + j * SS_THREADS;
      if (c < Hd) {
        const float y = fmaxf((v[j] - mean) * rstd * gv[j] + bv[j], 0.f);
        sm.arow[c] = __float2bfloat16_rn(y);
        sm.nrm[c] = y;
      }
    }
    __syncthreads();
  };
  auto put_vec = [&](float* dst) {
    return [dst](int, int n, const float (&v)[8]) {
      float4* d = reinterpret_cast<float4*>(dst + n);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    };
  };
  ss_nsplit_mma(sm, w, sm.xs, 0, SS_F, p.shadow + g[MMER_G_C0_W], nC0, nC1);
  ss_nsplit_reduce(sm, SS_F, sm.sv + SV_C0B, nC0, nC1, 1, false, put_vec(sc + SF_H1));
  ss_cluster_arrive();
  ss_nsplit_prefetch(w, p.shadow + g[MMER_G_C4_W], Hd, nC0, nC1);
  ss_cluster_wait();
  stp.mark();
  const bool hn_staged = Hd <= SS_HN;
  head_norm(sc + SF_H1, hn_staged ? sm.sv + SV_HN : p.params + g[MMER_G_C1_W],
            hn_staged ? sm.sv + SV_HN + SS_HN : p.params + g[MMER_G_C1_B]);
  ss_nsplit_mma(sm, w, sm.arow, 0, Hd, p.shadow + g[MMER_G_C4_W], nC0, nC1);
  ss_nsplit_reduce(sm, Hd, sm.sv + SV_C4B, nC0, nC1, 1, false, put_vec(sc + SF_H2));
  ss_cluster_arrive();
  // the output layer (fp32 masters, one class per warp): its first 512 columns fly during the last barrier
  const float* W8 = p.params + g[MMER_G_C8_W];
  float w8[16];
  float b8 = 0.f;
  if (rank == 0 && warp < p.classes) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w8[i] = (i * 32 + lane < Hd) ? __ldg(W8 + (long long)warp * Hd + i * 32 + lane) : 0.f;
    b8 = __ldg(p.params + g[MMER_G_C8_B] + warp);
  }
  ss_cluster_wait();
  stp.mark();
  if (rank == 0) {
    head_norm(sc + SF_H2, hn_staged ? sm.sv + SV_HN + 2 * SS_HN : p.params + g[MMER_G_C5_W],
              hn_staged ? sm.sv + SV_HN + 3 * SS_HN : p.params + g[MMER_G_C5_B]);
    const float* h2 = sm.nrm;
    if (warp < p.classes) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 32 + lane < Hd) a = fmaf(h2[i * 32 + lane], w8[i], a);
      for (int k = 512 + lane; k < Hd; k += 32) a = fmaf(h2[k], __ldg(W8 + (long long)warp * Hd + k), a);
      a = warp_sum(a);
      if (lane == 0) sm.red[32 + warp] = a + b8;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float* lg = sm.red + 32;
      float mx = -INFINITY;
      for (int c = 0; c < p.classes; ++c) mx = fmaxf(mx, lg[c]);
      float den = 0.f;
      for (int c = 0; c < p.classes; ++c) den += expf(lg[c] - mx);
      for (int c = 0; c < p.classes; ++c) {
        p.logits[c] = lg[c];
        p.probs[c] = expf(lg[c] - mx) / den;
      }
    }
  }
  stp.mark();
}

// row-major W[N][K] -> fragment order (ss_wp): one 16-byte chunk per thread
__global__ void serve_pack_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int K) {
  const long long chunks = (long long)N * K / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long block = i >> 6;                       // 64 chunks per (tile, k-step) block
    const int in = (int)(i & 63), half = in >> 5, L = in & 31, g = L >> 2, t = L & 3;
    const int kst = K >> 5;
    const int tile = (int)(block / kst), ks = (int)(block - (long long)tile * kst);
    const uint4 v = *reinterpret_cast<const uint4*>(src + (long long)(tile * 16 + half * 8 + g) * K + ks * 32 + t * 8);
    *reinterpret_cast<uint4*>(dst + i * 8) = v;
  }
}

}  // namespace

// returns 1 when this kernel does not apply (the caller falls back to serve.cu), 0 on success, < 0 on error
int serve_forward_small(const mmer_model* m, const void* packed, float* scratch, long long* stamps, int64_t scratch_floats_before_stamps,
                        cudaStream_t st) {
  if (packed == nullptr || m->T + 1 > 8 || m->layers > SS_MAXL || m->fused != SS_F || m->heads != SS_HEADS || m->ffn != SS_FFN || m->video_dim > SS_KX || m->audio_dim > SS_KX ||
      m->hidden > SS_MAXH || m->hidden % 256 != 0 || m->video_dim % 256 != 0 || m->audio_dim % 256 != 0 ||
      SF_ACC + 2LL * m->layers * SF_ACC_SZ > scratch_floats_before_stamps)
    return 1;
  auto kern = serve_small_kernel;
  static unsigned long long attr_done = 0ull;
  static unsigned long long ok16 = 0ull;
  const size_t smem = sizeof(SmemS);
  int dev = 0;
  cudaGetDevice(&dev);
  if (needs_func_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(serve small smem)");
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(SS_NC);
      cfg.blockDim = dim3(SS_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = SS_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1 && dev < 64) ok16 |= 1ull << dev;
    }
    (void)cudaGetLastError();
  }
  if (dev >= 64 || !((ok16 >> dev) & 1ull)) return 1;
  ServeParamsS p;
  p.T = m->T; p.S = m->T + 1;
  p.video_dim = m->video_dim; p.audio_dim = m->audio_dim; p.hidden = m->hidden; p.classes = m->classes; p.layers = m->layers;
  p.shadow = reinterpret_cast<const bf16*>(packed);
  p.params = m->params;
  for (int i = 0; i < MMER_G_COUNT; ++i) p.off_g[i] = m->off_g[i];
  for (int l = 0; l < SS_MAXL; ++l)
    for (int i = 0; i < MMER_L_COUNT; ++i) p.off_l[l][i] = m->off_l[l][i];
  p.fine = g_debug[MMER_DEBUG_SERVE_STAMPS];
  p.video = reinterpret_cast<const bf16*>(m->video);
  p.audio = reinterpret_cast<const bf16*>(m->audio);
  p.mask = m->has_mask ? m->mask : nullptr;
  p.scratch = scratch;
  p.stamps = stamps;
  p.logits = m->logits; p.probs = m->probs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(SS_NC);
  cfg.blockDim = dim3(SS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SS_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(serve_small)");
  MMER_LAUNCH_CHECK("serve_small_kernel");
  return 0;
}

// every matrix serve_small_kernel multiplies, from the row-major bf16 shadow into `packed` at the same offsets
int serve_pack_weights(const mmer_model* m, void* packed, cudaStream_t st) {
  const bf16* src = reinterpret_cast<const bf16*>(m->shadow);
  bf16* dst = reinterpret_cast<bf16*>(packed);
  auto one = [&](int64_t off, int N, int K) -> int {
    if (off < 0 || N % 16 != 0 || K % 32 != 0) return 0;     // not packable: serve_forward_small rejects such models
    const long long chunks = (long long)N * K / 8;
    const int blocks = (int)std::min<long long>((chunks + 255) / 256, 1184);
    serve_pack_kernel<<<blocks, 256, 0, st>>>(src + off, dst + off, N, K);
    MMER_LAUNCH_CHECK("serve_pack_kernel");
    return 0;
  };
  MMER_TRY(one(m->off_g[MMER_G_WV], m->fused, m->video_dim));
  MMER_TRY(one(m->off_g[MMER_G_WA], m->fused, m->audio_dim));
  for (int l = 0; l < m->layers; ++l) {
    MMER_TRY(one(m->off_l[l][MMER_L_IN_W], 3 * m->fused, m->fused));
    MMER_TRY(one(m->off_l[l][MMER_L_OUT_W], m->fused, m->fused));
    MMER_TRY(one(m->off_l[l][MMER_L_FF1_W], m->ffn, m->fused));
    MMER_TRY(one(m->off_l[l][MMER_L_FF2_W], m->fused, m->ffn));
  }
  MMER_TRY(one(m->off_g[MMER_G_C0_W], m->hidden, m->fused));
  MMER_TRY(one(m->off_g[MMER_G_C4_W], m->hidden, m->hidden));
  return 0;
}

}  // namespace mmer
