// Long-sequence attention on tensor cores (bf16, head size 64, 32 < S = T+1 <= 384): the T = 256 interpretability
// configuration of train2.py (BASELINE.json configs[3]; SDPA inside nn.MultiheadAttention, train2.py:173-176).
//
// One CTA per (sample, head), 8 warps.  Q, K, V (and dO in backward) of the head are [S][64] bf16 = S rows of 128 B:
// they arrive as 2-D TMA boxes (64 rows + a tail box) into 128B-swizzled shared-memory tiles, and every product is a
// warp-level m16n8k16 MMA on ldmatrix fragments of those tiles.  Nothing of size S x S is ever materialised:
//   forward   a warp owns 16 query rows at a time; pass 1 walks the key blocks for the row maxima and sums (online),
//             pass 2 recomputes the scores, normalises, (optionally writes the probabilities: return_attn), applies
//             dropout and accumulates O = P V
//   backward  pass A (per query tile): forward recomputation -> row statistics m, 1/l and D = dO . O in shared memory;
//             pass B (per query tile): dS = P o (dP o f - D) / sqrt(d), dQ = dS K;
//             pass C (per key tile, transposed domain so that no cross-warp reduction is needed):
//             S^T = K Q^T, dP^T = V dO^T, dV = Pd^T dO, dK = dS^T Q
// Scores are recomputed in every pass: attention is 8 % of the layer's FLOPs at T = 256 and the tensor pipe has room.
// Dropout decisions are regenerated from (seed, site, row, key) exactly as in the other attention kernels.
#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

static constexpr int AL_WARPS = 8;        // forward kernel (2 CTAs per SM)
static constexpr int AL_BWD_MAX_WARPS = 9;   // backward: 222 registers per thread allow 9 warps (288 threads)
static constexpr int AL_D = 64;

struct SwzRow {   // byte offset of (row, col) in a tile of 128-byte rows, 128B-swizzled (what TMA writes)
  __device__ __forceinline__ uint32_t operator()(int row, int col) const {
    return (uint32_t)row * 128u + (((((uint32_t)col >> 3) ^ (uint32_t)row) & 7u) << 4) + ((uint32_t)col & 7u) * 2u;
  }
};

// Per-lane byte offsets of the ldmatrix row addresses inside a swizzled tile.  Tile rows handled together start at
// multiples of 8, so the swizzle term (chunk ^ row) & 7 depends on the lane only: computed once per kernel, the address
// of every ldmatrix is base + immediate.
struct LaneOff {
  uint32_t nt[2];   // B fragments of 8 tile rows (non-transposed), k2 = 0, 1: row (lane & 7), 16-byte chunk k2*4 + (lane >> 3)
  uint32_t tr[4];   // 16 tile rows, chunk pair n2 = 0..3: row (lane & 7) + 8 * ((lane >> 3) & 1), chunk n2*2 + (lane >> 4)
};
__device__ __forceinline__ LaneOff lane_offsets(int lane) {
  LaneOff o;
  const uint32_t r7 = lane & 7;
#pragma unroll
  for (int k2 = 0; k2 < 2; ++k2) o.nt[k2] = r7 * 128u + ((((uint32_t)(k2 * 4 + (lane >> 3)) ^ r7) & 7u) << 4);
  const uint32_t lrow = r7 + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) o.tr[n2] = lrow * 128u + ((((uint32_t)(n2 * 2 + (lane >> 4)) ^ r7) & 7u) << 4);
  return o;
}
// A fragments (16 rows x 64 columns = 4 k-steps) of rows [r0, r0+16) of a tile (r0 % 8 == 0)
__device__ __forceinline__ void load_a16(uint32_t tile_a, int r0, const LaneOff& lo, uint32_t (&a)[4][4]) {
  const uint32_t base = tile_a + (uint32_t)r0 * 128u;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(base + lo.tr[ks], a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}
// c[nt][..] = A(16 x 64) . X[rows c0 .. c0+32)^T : 16 x 32 block of products against 32 rows of tile X (c0 % 8 == 0)
__device__ __forceinline__ void block_nt(const uint32_t (&a)[4][4], uint32_t x_a, int c0, const LaneOff& lo, float (&c)[4][4]) {
  const uint32_t b0 = x_a + (uint32_t)c0 * 128u + lo.nt[0], b1 = x_a + (uint32_t)c0 * 128u + lo.nt[1];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    uint32_t xf[4][2];
    ldsm_x4(b0 + nt * 1024, xf[0][0], xf[0][1], xf[1][0], xf[1][1]);
    ldsm_x4(b1 + nt * 1024, xf[2][0], xf[2][1], xf[3][0], xf[3][1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) c[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_bf16_16816(c[nt], a[ks][0], a[ks][1], a[ks][2], a[ks][3], xf[ks][0], xf[ks][1]);
  }
}
// acc(16 x 64) += P(16 x 32, as two A fragments) . X[rows r0 .. r0+32) (X row-major: ldmatrix.trans; r0 % 8 == 0)
__device__ __forceinline__ void acc_rows(float (&acc)[8][4], const uint32_t (&pa)[2][4], uint32_t x_a, int r0, const LaneOff& lo) {
  const uint32_t base = x_a + (uint32_t)r0 * 128u;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(base + ks * 2048 + lo.tr[n2], b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * n2], pa[ks][0], pa[ks][1], pa[ks][2], pa[ks][3], b0, b1);
      mma_bf16_16816(acc[2 * n2 + 1], pa[ks][0], pa[ks][1], pa[ks][2], pa[ks][3], b2, b3);
    }
}
__device__ __forceinline__ void pack_block(const float (&c)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}
// 16 x 64 accumulator tile -> bf16 rows of a (non-swizzled) per-warp staging tile [16][64], then 16-byte coalesced
// stores of the rows < nrows to global (row stride ld elements)
__device__ __forceinline__ void store_tile_global(const float (&acc)[8][4], uint8_t* stage, bf16* __restrict__ dst, long long ld,
                                                  int nrows, int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
      *reinterpret_cast<uint32_t*>(stage + (g + 8 * r) * 144 + (nd * 8 + t * 2) * 2) = pack_bf16x2(acc[nd][2 * r], acc[nd][2 * r + 1]);
  __syncwarp();
  for (int idx = lane; idx < 16 * 8; idx += 32) {
    const int r = idx >> 3, c8 = idx & 7;
    if (r < nrows) *reinterpret_cast<uint4*>(dst + (long long)r * ld + c8 * 8) = *reinterpret_cast<const uint4*>(stage + r * 144 + c8 * 16);
  }
}

struct LongGeom {
  int B, Tn, S, SR, H, F;        // SR = S rounded up to 8: the rows TMA delivers per tile
  int nfull, tail;               // 64-row TMA boxes and the rows of the tail box (multiple of 8, may be 0)
  uint32_t tile_bytes;           // tile allocation: S rounded up to 32 rows of 128 B (rows >= SR are zero-filled)
  uint32_t load_bytes;           // SR * 128
};

// loads tile `which` (0 Q, 1 K, 2 V of the packed qkv matrix; 3 = dO) of (b, h)
__device__ __forceinline__ void load_head_tile(uint32_t dst, const CUtensorMap* m64, const CUtensorMap* mtail, uint32_t bar,
                                               const LongGeom& g, int col, long long row0) {
  for (int c = 0; c < g.nfull; ++c) tma_load_2d(dst + c * 8192, m64, bar, col, (int)(row0 + c * 64));
  if (g.tail > 0) tma_load_2d(dst + g.nfull * 8192, mtail, bar, col, (int)(row0 + g.nfull * 64));
}

// rows [SR, tile rows) of `ntiles` consecutive tiles are never written by TMA: zero them (they are multiplied by
// exactly-zero probabilities, which must not meet NaN bit patterns)
__device__ __forceinline__ void zero_pad_rows(uint8_t* tiles, int ntiles, const LongGeom& g) {
  const uint32_t pad = g.tile_bytes - g.load_bytes;
  for (int i = threadIdx.x * 16; i < (int)(ntiles * pad); i += blockDim.x * 16) {
    const int tl = i / (int)pad, o = i % (int)pad;
    *reinterpret_cast<uint4*>(tiles + (size_t)tl * g.tile_bytes + g.load_bytes + o) = make_uint4(0u, 0u, 0u, 0u);
  }
}
// valid[j] (bytes) and vbits[j / 32] (bit j % 32): key j takes part (j < S and not padded).  blockDim.x % 32 == 0.
__device__ __forceinline__ void fill_valid(uint8_t* valid, uint32_t* vbits, const uint8_t* __restrict__ mask, int b, const LongGeom& g) {
  const int n = (g.SR + 32 + 31) & ~31;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    bool ok = j < g.S;
    if (ok && j < g.Tn && mask != nullptr) ok = mask[(long long)b * g.Tn + j] == 0;
    if (j < g.SR + 32) valid[j] = ok ? 1 : 0;
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) vbits[j >> 5] = w;
  }
}
constexpr int AL_VWORDS = 16;   // (384 + 32 + 31) / 32 bit words, rounded up

// online row statistics of one 16 x 32 score block (rows g and g+8 of the tile; quad-uniform maxima)
__device__ __forceinline__ void stats_update(const float (&sc)[4][4], const uint8_t* valid, int key0, int t, float sl2,
                                             float (&m)[2], float (&l)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float bm = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (valid[key0 + nt * 8 + t * 2 + e]) bm = fmaxf(bm, sc[nt][r * 2 + e]);
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
    const float mn = fmaxf(m[r], bm);
    if (mn == -INFINITY) continue;            // nothing valid yet in this row
    float add = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (valid[key0 + nt * 8 + t * 2 + e]) add += ex2_approx((sc[nt][r * 2 + e] - mn) * sl2);
    l[r] = l[r] * ex2_approx((m[r] - mn) * sl2) + add;    // l is this lane's partial sum; m is shared by the quad
    m[r] = mn;
  }
}

// bytes of the validity region: valid[SR + 32] bytes, then AL_VWORDS bit words
__device__ __host__ __forceinline__ int valid_region_bytes(int SR) { return ((SR + 32 + 15) & ~15) + AL_VWORDS * 4; }

// One 16 x 32 score block of the one-pass forward: running maxima (O and the row sums are rescaled only when some row
// of the warp's tile raises its maximum), unnormalised probabilities back into sc (dropout applied), row sums into l.
// vm: this lane's validity bits (bit nt * 8 + e); ALL: every key of the block takes part.  keep: this lane's dropout
// keep bits of the block (bit nt * 8 + t * 2 + e), handed to the backward kernel through global memory.
template <bool DROP, bool ALL>
__device__ __forceinline__ void fwd_block_softmax(float (&sc)[4][4], uint32_t vm, float sl2, float (&m)[2], float (&l)[2],
                                                  float (&o)[8][4], const DropCfg& dc, const uint32_t (&oct)[2], uint32_t mult, int t,
                                                  uint32_t (&keep)[2]) {
  float mn[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float bm = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (ALL) {
        bm = fmaxf(bm, fmaxf(sc[nt][r * 2], sc[nt][r * 2 + 1]));
      } else {
        bm = fmaxf(bm, ((vm >> (nt * 8)) & 1u) ? sc[nt][r * 2] : -INFINITY);
        bm = fmaxf(bm, ((vm >> (nt * 8 + 1)) & 1u) ? sc[nt][r * 2 + 1] : -INFINITY);
      }
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
    mn[r] = fmaxf(m[r], bm);
  }
  if (__any_sync(0xffffffffu, mn[0] != m[0] || mn[1] != m[1])) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float ms = (mn[r] == -INFINITY) ? 0.f : mn[r] * sl2;
      const float corr = ex2_approx(fmaf(m[r], sl2, -ms));       // m = -inf -> 0; unchanged maximum -> 1
      l[r] *= corr;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) { o[nd][2 * r] *= corr; o[nd][2 * r + 1] *= corr; }
      m[r] = mn[r];
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float ms = (m[r] == -INFINITY) ? 0.f : m[r] * sl2;     // nothing valid yet: every p below is selected to 0
    float add = 0.f;
    uint32_t kb = 0u;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p0 = ex2_approx(fmaf(sc[nt][r * 2], sl2, -ms));
      float p1 = ex2_approx(fmaf(sc[nt][r * 2 + 1], sl2, -ms));
      if (!ALL) {
        p0 = ((vm >> (nt * 8)) & 1u) ? p0 : 0.f;
        p1 = ((vm >> (nt * 8 + 1)) & 1u) ? p1 : 0.f;
      }
      add += p0 + p1;
      if (DROP) {
        // the lane's pair nt sits in octet oct[r] + nt of the row, word t (mult = drop_mult(t))
        const uint32_t w = drop_word(drop_base(dc, oct[r] + nt), mult);
        const bool k0 = (w << 16) >= dc.thr_hi, k1 = w >= dc.thr_hi;
        kb |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (nt * 8);
        p0 = k0 ? p0 * dc.scale : 0.f;
        p1 = k1 ? p1 * dc.scale : 0.f;
      }
      sc[nt][r * 2] = p0;
      sc[nt][r * 2 + 1] = p1;
    }
    l[r] += add;
    keep[r] = kb << (t * 2);
  }
}
// the four lanes of a quad hold the keep bits of one (row, 32-key block): OR them, lane t == 0 stores the word
__device__ __forceinline__ void store_keep_word(uint32_t w, int t, bool row_ok, uint32_t* dst) {
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  if (t == 0 && row_ok) *dst = w;
}

// PROBS (return_attn): two passes over the key blocks (statistics, then normalised probabilities written out and P V).
// Otherwise ONE pass with running maxima: O and the row sums are rescaled when a block raises the maximum, the division
// by the row sum happens once at the end (a third fewer MMAs, every exponential computed once).
template <bool DROP, bool PROBS>
__global__ void __launch_bounds__(AL_WARPS * 32, 2)
mha_fwd_long_kernel(const __grid_constant__ CUtensorMap m64, const __grid_constant__ CUtensorMap mtail,
                    const uint8_t* __restrict__ mask, bf16* __restrict__ out, float* __restrict__ probs,
                    float2* __restrict__ lse, uint32_t* __restrict__ keep_g, LongGeom g, DropCfg dc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int S = g.S, H = g.H, F = g.F;
  const long long bh = blockIdx.x;
  const int b = (int)(bh / H), h = (int)(bh % H);
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + g.tile_bytes;
  uint8_t* v_s = k_s + g.tile_bytes;
  uint8_t* valid = v_s + g.tile_bytes;
  uint32_t* vbits = reinterpret_cast<uint32_t*>(valid + ((g.SR + 32 + 15) & ~15));
  uint64_t* bar = reinterpret_cast<uint64_t*>(valid + valid_region_bytes(g.SR));
  const uint32_t bar_a = smem_u32(bar), q_a = smem_u32(q_s), k_a = smem_u32(k_s), v_a = smem_u32(v_s);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  fill_valid(valid, vbits, mask, b, g);
  zero_pad_rows(q_s, 3, g);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_a, 3u * g.load_bytes);
    const long long row0 = (long long)b * S;
    load_head_tile(q_a, &m64, &mtail, bar_a, g, h * AL_D, row0);
    load_head_tile(k_a, &m64, &mtail, bar_a, g, F + h * AL_D, row0);
    load_head_tile(v_a, &m64, &mtail, bar_a, g, 2 * F + h * AL_D, row0);
  }
  mbar_wait(bar_a, 0);
  const float sl2 = rsqrtf((float)AL_D) * 1.4426950408889634f;
  const int dstride = att_drop_stride(S);
  const int nkb = (S + 31) / 32;
  const LaneOff lo = lane_offsets(lane);
  for (int q0 = warp * 16; q0 < S; q0 += AL_WARPS * 16) {
    uint32_t qa[4][4];
    load_a16(q_a, q0, lo, qa);
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float inv[2];
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[nd][i] = 0.f;
    if (PROBS) {
      for (int kb = 0; kb < nkb; ++kb) {
        float sc[4][4];
        block_nt(qa, k_a, kb * 32, lo, sc);
        stats_update(sc, valid, kb * 32, t, sl2, m, l);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float s = l[r];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        inv[r] = 1.f / s;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        float sc[4][4];
        block_nt(qa, k_a, kb * 32, lo, sc);
        uint32_t kw[2] = {0u, 0u};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = q0 + gq + 8 * r;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int j = kb * 32 + nt * 8 + t * 2;
            float p0 = valid[j] ? ex2_approx((sc[nt][r * 2] - m[r]) * sl2) * inv[r] : 0.f;
            float p1 = valid[j + 1] ? ex2_approx((sc[nt][r * 2 + 1] - m[r]) * sl2) * inv[r] : 0.f;
            if (i < S) {
              if (probs != nullptr) {
                if (j < S) probs[(bh * S + i) * S + j] = p0;
                if (j + 1 < S) probs[(bh * S + i) * S + j + 1] = p1;
              }
              if (DROP) {
                float f0, f1;
                drop2(dc, att_drop_index(bh * S + i, j, dstride), f0, f1);
                p0 *= f0;
                p1 *= f1;
                kw[r] |= ((f0 != 0.f ? 1u : 0u) | (f1 != 0.f ? 2u : 0u)) << (nt * 8 + t * 2);
              }
            }
            sc[nt][r * 2] = p0;
            sc[nt][r * 2 + 1] = p1;
          }
        }
        if (DROP && keep_g != nullptr) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
            store_keep_word(kw[r], t, q0 + gq + 8 * r < S, keep_g + (bh * S + q0 + gq + 8 * r) * nkb + kb);
        }
        uint32_t pa[2][4];
        pack_block(sc, pa);
        acc_rows(o, pa, v_a, kb * 32, lo);
      }
    } else {
      // dropout index of (row, key 0) per fragment row; rows >= S (last tile only) produce unused output rows
      const uint32_t orow[2] = {(uint32_t)((bh * S + q0 + gq) * (dstride >> 3)), (uint32_t)((bh * S + q0 + gq + 8) * (dstride >> 3))};
      const uint32_t mult_t = drop_mult(t);
      for (int kb = 0; kb < nkb; ++kb) {
        float sc[4][4];
        block_nt(qa, k_a, kb * 32, lo, sc);
        const uint32_t vw = vbits[kb];
        const uint32_t oct[2] = {orow[0] + kb * 4, orow[1] + kb * 4};
        uint32_t keep[2];
        if (vw == 0xffffffffu)   // warp-uniform: every block but the last, unless a padding mask is given
          fwd_block_softmax<DROP, true>(sc, 0u, sl2, m, l, o, dc, oct, mult_t, t, keep);
        else
          fwd_block_softmax<DROP, false>(sc, vw >> (t * 2), sl2, m, l, o, dc, oct, mult_t, t, keep);
        if (DROP && keep_g != nullptr) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
            store_keep_word(keep[r], t, q0 + gq + 8 * r < S, keep_g + (bh * S + q0 + gq + 8 * r) * nkb + kb);
        }
        uint32_t pa[2][4];
        pack_block(sc, pa);
        acc_rows(o, pa, v_a, kb * 32, lo);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float sum = l[r];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv[r] = 1.f / sum;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) { o[nd][2 * r] *= inv[r]; o[nd][2 * r + 1] *= inv[r]; }
      }
    }
    // softmax statistics of the rows for the backward kernel (saves it two of its four score recomputations)
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (lse != nullptr && t == 0 && q0 + gq + 8 * r < S) lse[bh * S + q0 + gq + 8 * r] = make_float2(m[r], inv[r]);
    // O overwrites this warp's own (now dead) Q rows, then leaves as 16-byte coalesced row stores
    const SwzRow off;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        *reinterpret_cast<uint32_t*>(q_s + off(q0 + gq + 8 * r, nd * 8 + t * 2)) = pack_bf16x2(o[nd][2 * r], o[nd][2 * r + 1]);
    __syncwarp();
    for (int idx = lane; idx < 16 * 8; idx += 32) {
      const int r = idx >> 3, c8 = idx & 7;
      if (q0 + r < S)
        *reinterpret_cast<uint4*>(out + ((long long)b * S + q0 + r) * F + h * AL_D + c8 * 8) =
            *reinterpret_cast<const uint4*>(q_s + off(q0 + r, c8 * 8));
    }
  }
}

// Backward, query-major block: sc <- dS = [ex2(s * sl2 - m * sl2) * (scale / l)] * (dP * f - D) for a 16 x 32 block
// (rows beyond the sequence carry scale / l = 0).  vm: this lane's validity bits.  Dropout: oct[r] = octet index of
// (row, key 0 of the block) -- the lane's pair nt sits in octet oct[r] + nt, word t (mult hoisted).  keep[r] collects
// this lane's keep bits of the block (bit nt * 8 + t * 2 + e) for the transposed pass.
// GB: the keep bits of the block come from the forward pass (keep[r] holds the row's word on entry) -- no hash.
template <bool DROP, bool ALL, bool GB>
__device__ __forceinline__ void bwd_block_ds(float (&sc)[4][4], const float (&dp)[4][4], uint32_t vm, float sl2, const float (&ms)[2],
                                             const float (&is)[2], const float (&dsum)[2], const DropCfg& dc,
                                             const uint32_t (&oct)[2], uint32_t mult, int t, uint32_t (&keep)[2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    uint32_t kb = 0u;
    const uint32_t given = keep[r] >> (t * 2);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p0 = ex2_approx(fmaf(sc[nt][r * 2], sl2, -ms[r])) * is[r];
      float p1 = ex2_approx(fmaf(sc[nt][r * 2 + 1], sl2, -ms[r])) * is[r];
      if (!ALL) {
        p0 = ((vm >> (nt * 8)) & 1u) ? p0 : 0.f;
        p1 = ((vm >> (nt * 8 + 1)) & 1u) ? p1 : 0.f;
      }
      float g0 = dp[nt][r * 2] - dsum[r], g1 = dp[nt][r * 2 + 1] - dsum[r];
      if (DROP) {
        bool k0, k1;
        if (GB) {
          k0 = (given >> (nt * 8)) & 1u;
          k1 = (given >> (nt * 8 + 1)) & 1u;
        } else {
          const uint32_t w = drop_word(drop_base(dc, oct[r] + nt), mult);
          k0 = (w << 16) >= dc.thr_hi;
          k1 = w >= dc.thr_hi;
          kb |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (nt * 8);
        }
        g0 = fmaf(dp[nt][r * 2], k0 ? dc.scale : 0.f, -dsum[r]);
        g1 = fmaf(dp[nt][r * 2 + 1], k1 ? dc.scale : 0.f, -dsum[r]);
      }
      sc[nt][r * 2] = p0 * g0;
      sc[nt][r * 2 + 1] = p1 * g1;
    }
    if (!GB) keep[r] = kb << (t * 2);
  }
}
// Backward, key-major (transposed) block: rows = two keys of this lane (jj), columns = queries.  pd <- P o f (for dV),
// st <- dS^T (for dK).  qs: per-query statistics of this lane's first column (m * sl2, 1 / l, D, scale / l).
// Dropout factor of element (query i, key j): word (j >> 1) & 3 of octet (bh * S + i) * stride / 8 + j / 8, low or high
// half by j & 1 -- the key part is fixed per row (dmul, dsh), the query part advances by doct per column.
// BITS: the keep decisions come from the bit map the query-major pass wrote (kbits: row of this lane's first column, nw
// words per query row, word = key block of this tile, kpos[r] = bit of key jj[r]) instead of a hash per element.
template <bool DROP, bool ALL, bool BITS>
__device__ __forceinline__ void bwd_block_dst(float (&st)[4][4], const float (&dpt)[4][4], float (&pd)[4][4], const float4* qs,
                                              const bool (&jv)[2], float sl2, const DropCfg& dc, long long oct0, long long doct,
                                              const int (&jj)[2], const uint32_t (&dmul)[2], const int (&dsh)[2],
                                              const uint32_t* kbits, int nw, const int (&kpos)[2]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float4 qs0 = qs[nt * 8], qs1 = qs[nt * 8 + 1];
    uint32_t kw0 = 0u, kw1 = 0u;
    if (DROP && BITS) {
      kw0 = kbits[(nt * 8) * nw];
      kw1 = kbits[(nt * 8 + 1) * nw];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float e0 = ex2_approx(fmaf(st[nt][r * 2], sl2, -qs0.x)), e1 = ex2_approx(fmaf(st[nt][r * 2 + 1], sl2, -qs1.x));
      float p0 = e0 * qs0.y, p1 = e1 * qs1.y;                  // P (for dV)
      float s0 = e0 * qs0.w, s1 = e1 * qs1.w;                  // P * scale (for dS)
      if (!ALL) {
        p0 = jv[r] ? p0 : 0.f; p1 = jv[r] ? p1 : 0.f;
        s0 = jv[r] ? s0 : 0.f; s1 = jv[r] ? s1 : 0.f;
      }
      float g0 = dpt[nt][r * 2] - qs0.z, g1 = dpt[nt][r * 2 + 1] - qs1.z;
      if (DROP) {
        float f0, f1;
        if (BITS) {
          f0 = ((kw0 >> kpos[r]) & 1u) ? dc.scale : 0.f;
          f1 = ((kw1 >> kpos[r]) & 1u) ? dc.scale : 0.f;
        } else {
          const long long o0 = oct0 + nt * 8 * doct + (jj[r] >> 3);
          const uint32_t w0 = drop_word(drop_base(dc, (uint32_t)o0), dmul[r]);
          const uint32_t w1 = drop_word(drop_base(dc, (uint32_t)(o0 + doct)), dmul[r]);
          f0 = (w0 << dsh[r]) >= dc.thr_hi ? dc.scale : 0.f;
          f1 = (w1 << dsh[r]) >= dc.thr_hi ? dc.scale : 0.f;
        }
        p0 *= f0;
        p1 *= f1;
        g0 = fmaf(dpt[nt][r * 2], f0, -qs0.z);
        g1 = fmaf(dpt[nt][r * 2 + 1], f1, -qs1.z);
      }
      pd[nt][r * 2] = p0;
      pd[nt][r * 2 + 1] = p1;
      st[nt][r * 2] = s0 * g0;
      st[nt][r * 2 + 1] = s1 * g1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// The backward kernel's warp count is a launch parameter: a warp owns 16-row query / key tiles, and S = 257 has 17 of
// them -- 9 warps need two rounds per pass where 8 need three (17 / 8 = 2.1).
// BITS: room for the dropout keep-bit map [query][key block] in shared memory (always at S = 257; not at S = 384)
// GB (implies BITS): the forward pass wrote the map to global memory (keep_g); it is copied in at the start and no
// element of this kernel is hashed.
template <bool DROP, bool BITS, bool GB>
__global__ void __launch_bounds__(AL_BWD_MAX_WARPS * 32, 1)
mha_bwd_long_kernel(const __grid_constant__ CUtensorMap m64, const __grid_constant__ CUtensorMap mtail,
                    const __grid_constant__ CUtensorMap d64, const __grid_constant__ CUtensorMap dtail,
                    const uint8_t* __restrict__ mask, bf16* __restrict__ dqkv, const float2* __restrict__ lse,
                    const bf16* __restrict__ fwd_out, const uint32_t* __restrict__ keep_g, LongGeom g, DropCfg dc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int S = g.S, H = g.H, F = g.F;
  const long long bh = blockIdx.x;
  const int b = (int)(bh / H), h = (int)(bh % H);
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + g.tile_bytes;
  uint8_t* v_s = k_s + g.tile_bytes;
  uint8_t* do_s = v_s + g.tile_bytes;
  float4* st_q = reinterpret_cast<float4*>(do_s + g.tile_bytes);   // per query: (m * sl2, 1 / l, D = dO . O, scale / l)  [SR + 32]
  uint8_t* valid = reinterpret_cast<uint8_t*>(st_q + g.SR + 32);
  uint32_t* vbits = reinterpret_cast<uint32_t*>(valid + ((g.SR + 32 + 15) & ~15));
  uint8_t* stage = valid + valid_region_bytes(g.SR) + warp * (16 * 144);
  const int nwarps = (int)(blockDim.x >> 5);
  uint64_t* bar = reinterpret_cast<uint64_t*>(valid + valid_region_bytes(g.SR) + AL_BWD_MAX_WARPS * 16 * 144);
  uint32_t* kbits = reinterpret_cast<uint32_t*>(bar + 2);             // [SR + 32][nkb] keep bits (DROP && BITS)
  const uint32_t bar_a = smem_u32(bar), q_a = smem_u32(q_s), k_a = smem_u32(k_s), v_a = smem_u32(v_s), do_a = smem_u32(do_s);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  fill_valid(valid, vbits, mask, b, g);
  zero_pad_rows(q_s, 4, g);
  for (int j = threadIdx.x; j < g.SR + 32; j += blockDim.x) st_q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (DROP && GB) {
    const int nw = (S + 31) / 32;
    const uint32_t* src = keep_g + bh * S * nw;
    for (int j = threadIdx.x; j < (g.SR + 32) * nw; j += blockDim.x) kbits[j] = j < S * nw ? src[j] : 0u;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_a, 4u * g.load_bytes);
    const long long row0 = (long long)b * S;
    load_head_tile(q_a, &m64, &mtail, bar_a, g, h * AL_D, row0);
    load_head_tile(k_a, &m64, &mtail, bar_a, g, F + h * AL_D, row0);
    load_head_tile(v_a, &m64, &mtail, bar_a, g, 2 * F + h * AL_D, row0);
    load_head_tile(do_a, &d64, &dtail, bar_a, g, h * AL_D, row0);
  }
  mbar_wait(bar_a, 0);
  const float scale = rsqrtf((float)AL_D);
  const float sl2 = scale * 1.4426950408889634f;
  const int dstride = att_drop_stride(S);
  const int nkb = (S + 31) / 32;
  const SwzRow off;
  const LaneOff lo = lane_offsets(lane);

  // ---------------- pass A + B per query tile: statistics, then dQ
  for (int q0 = warp * 16; q0 < S; q0 += nwarps * 16) {
    uint32_t qa[4][4], doa[4][4];
    load_a16(q_a, q0, lo, qa);
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float inv[2];
    float dsum[2] = {0.f, 0.f};
    load_a16(do_a, q0, lo, doa);
    if (lse != nullptr) {
      // The forward kernel stored (max, 1 / sum) of every row, and D_i = sum_j P_ij f_ij dP_ij equals dO_i . O_i with
      // the forward output O (dropout included): two of the four score recomputations disappear.
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = q0 + gq + 8 * r;
        if (i < S) {
          const float2 st = lse[bh * S + i];
          m[r] = st.x;
          inv[r] = st.y;
          const bf16* orow = fwd_out + ((long long)b * S + i) * F + h * AL_D + t * 16;
          float ov[16], dv[16];
          load8(orow, *reinterpret_cast<float(*)[8]>(ov));
          load8(orow + 8, *reinterpret_cast<float(*)[8]>(ov + 8));
          load8(reinterpret_cast<const bf16*>(do_s + off(i, t * 16)), *reinterpret_cast<float(*)[8]>(dv));
          load8(reinterpret_cast<const bf16*>(do_s + off(i, t * 16 + 8)), *reinterpret_cast<float(*)[8]>(dv + 8));
#pragma unroll
          for (int c = 0; c < 16; ++c) dsum[r] = fmaf(ov[c], dv[c], dsum[r]);
        } else {
          m[r] = 0.f;
          inv[r] = 0.f;
        }
      }
    } else {
    for (int kb = 0; kb < nkb; ++kb) {
      float sc[4][4];
      block_nt(qa, k_a, kb * 32, lo, sc);
      stats_update(sc, valid, kb * 32, t, sl2, m, l);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float s = l[r];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      inv[r] = 1.f / s;
    }
    // D_i = sum_j P_ij f_ij dP_ij  with dP_ij = dO_i . V_j   (equals dO_i . O_i; accumulated block by block)
    for (int kb = 0; kb < nkb; ++kb) {
      float sc[4][4], dp[4][4];
      block_nt(qa, k_a, kb * 32, lo, sc);
      block_nt(doa, v_a, kb * 32, lo, dp);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = q0 + gq + 8 * r;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int j = kb * 32 + nt * 8 + t * 2;
          const float p0 = valid[j] ? ex2_approx((sc[nt][r * 2] - m[r]) * sl2) * inv[r] : 0.f;
          const float p1 = valid[j + 1] ? ex2_approx((sc[nt][r * 2 + 1] - m[r]) * sl2) * inv[r] : 0.f;
          float f0 = 1.f, f1 = 1.f;
          if (dc.thr && i < S) drop2(dc, att_drop_index(bh * S + i, j, dstride), f0, f1);
          dsum[r] = fmaf(p0 * f0, dp[nt][r * 2], dsum[r]);
          dsum[r] = fmaf(p1 * f1, dp[nt][r * 2 + 1], dsum[r]);
        }
      }
    }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
      const int i = q0 + gq + 8 * r;
      if (i >= S) { m[r] = 0.f; inv[r] = 0.f; }        // rows beyond the sequence: every probability below is exactly 0
      if (t == 0) st_q[i] = make_float4(m[r] * sl2, inv[r], dsum[r], inv[r] * scale);
    }
    // dQ = dS K
    float dq[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[nd][i] = 0.f;
    {
      // dS = P o (dP o f - D) / sqrt(d) = [ex2(s * sl2 - m * sl2) * (scale / l)] * (dP * f - D); rows >= S have 1 / l = 0
      const float ms[2] = {m[0] * sl2, m[1] * sl2}, is[2] = {inv[0] * scale, inv[1] * scale};
      const uint32_t orow[2] = {(uint32_t)((bh * S + q0 + gq) * (dstride >> 3)), (uint32_t)((bh * S + q0 + gq + 8) * (dstride >> 3))};
      const uint32_t mult_t = drop_mult(t);
      for (int kb = 0; kb < nkb; ++kb) {
        float sc[4][4], dp[4][4];
        block_nt(qa, k_a, kb * 32, lo, sc);
        block_nt(doa, v_a, kb * 32, lo, dp);
        const uint32_t vw = vbits[kb];
        const uint32_t oct[2] = {orow[0] + kb * 4, orow[1] + kb * 4};
        uint32_t keep[2] = {0u, 0u};
        if (DROP && GB) {
          keep[0] = kbits[(q0 + gq) * nkb + kb];
          keep[1] = kbits[(q0 + gq + 8) * nkb + kb];
        }
        if (vw == 0xffffffffu) bwd_block_ds<DROP, true, GB>(sc, dp, 0u, sl2, ms, is, dsum, dc, oct, mult_t, t, keep);
        else bwd_block_ds<DROP, false, GB>(sc, dp, vw >> (t * 2), sl2, ms, is, dsum, dc, oct, mult_t, t, keep);
        if (DROP && BITS && !GB) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            uint32_t w = keep[r];
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if (t == 0) kbits[(q0 + gq + 8 * r) * nkb + kb] = w;
          }
        }
        uint32_t dsa[2][4];
        pack_block(sc, dsa);
        acc_rows(dq, dsa, k_a, kb * 32, lo);
      }
    }
    store_tile_global(dq, stage, dqkv + ((long long)b * S + q0) * 3 * F + h * AL_D, 3 * F, S - q0, lane);
  }
  __syncthreads();   // row statistics of every query are in shared memory

  // ---------------- pass C per key tile (transposed domain): dV = Pd^T dO, dK = dS^T Q
  const int nqb = (S + 31) / 32;
  for (int k0 = warp * 16; k0 < S; k0 += nwarps * 16) {
    uint32_t ka[4][4], va[4][4];
    load_a16(k_a, k0, lo, ka);
    load_a16(v_a, k0, lo, va);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) { dk[nd][i] = 0.f; dv[nd][i] = 0.f; }
    // keys of this tile: validity per fragment row (uniform fast path when all 16 take part); dropout: the key part of
    // the element index and which 16-bit half of which word of the octet it selects are fixed per row
    const bool jv[2] = {valid[k0 + gq] != 0, valid[k0 + gq + 8] != 0};
    const bool tile_all = __all_sync(0xffffffffu, jv[0] && jv[1]);
    const int jj[2] = {k0 + gq, k0 + gq + 8};
    const uint32_t dmul[2] = {drop_mult((jj[0] >> 1) & 3), drop_mult((jj[1] >> 1) & 3)};
    const int dsh[2] = {(jj[0] & 1) ? 0 : 16, (jj[1] & 1) ? 0 : 16};
    const long long doct = (long long)dstride >> 3;                 // octets per (head, query) row
    const int kpos[2] = {(k0 & 31) + gq, (k0 & 31) + gq + 8};
    for (int qb = 0; qb < nqb; ++qb) {
      float st[4][4], dpt[4][4], pd[4][4];
      block_nt(ka, q_a, qb * 32, lo, st);     // S^T block: rows = keys k0.., columns = queries qb*32..
      block_nt(va, do_a, qb * 32, lo, dpt);   // dP^T block
      const uint32_t* kb0 = kbits + (qb * 32 + t * 2) * nkb + (k0 >> 5);
      if (tile_all)
        bwd_block_dst<DROP, true, BITS>(st, dpt, pd, st_q + qb * 32 + t * 2, jv, sl2, dc, (bh * S + qb * 32 + t * 2) * doct, doct, jj, dmul, dsh,
                                        kb0, nkb, kpos);
      else
        bwd_block_dst<DROP, false, BITS>(st, dpt, pd, st_q + qb * 32 + t * 2, jv, sl2, dc, (bh * S + qb * 32 + t * 2) * doct, doct, jj, dmul,
                                         dsh, kb0, nkb, kpos);
      uint32_t a2[2][4];
      pack_block(pd, a2);
      acc_rows(dv, a2, do_a, qb * 32, lo);
      pack_block(st, a2);
      acc_rows(dk, a2, q_a, qb * 32, lo);
    }
    bf16* base = dqkv + ((long long)b * S + k0) * 3 * F + h * AL_D;
    store_tile_global(dk, stage, base + F, 3 * F, S - k0, lane);
    store_tile_global(dv, stage, base + 2 * F, 3 * F, S - k0, lane);
  }
  (void)off;
}

// ---------------------------------------------------------------------------------------------------------
static int long_geom(int B, int Tn, int H, LongGeom* g) {
  g->B = B; g->Tn = Tn; g->S = Tn + 1; g->H = H; g->F = H * AL_D;
  g->SR = (g->S + 7) & ~7;
  g->nfull = g->SR / 64;
  g->tail = g->SR - g->nfull * 64;
  g->tile_bytes = (uint32_t)((g->S + 31) & ~31) * 128u;
  g->load_bytes = (uint32_t)g->SR * 128u;
  return 0;
}
static size_t long_smem(const LongGeom& g, int tiles, bool backward) {
  size_t n = (size_t)tiles * g.tile_bytes + valid_region_bytes(g.SR) + 16;
  if (backward) n += (size_t)4 * (g.SR + 32) * sizeof(float) + (size_t)AL_BWD_MAX_WARPS * 16 * 144;
  return n;
}
static int long_maps(const void* ptr, int cols_total, const LongGeom& g, CUtensorMap* m64, CUtensorMap* mtail) {
  const uint64_t rows = (uint64_t)g.B * g.S;
  MMER_TRY(make_tma_map_bf16(m64, ptr, (uint64_t)cols_total, rows, (uint64_t)cols_total, 64, g.nfull > 0 ? 64 : (uint32_t)g.tail));
  MMER_TRY(make_tma_map_bf16(mtail, ptr, (uint64_t)cols_total, rows, (uint64_t)cols_total, 64, g.tail > 0 ? (uint32_t)g.tail : 64));
  return 0;
}

bool mha_long_supported(int Tn, int d, int dtype) { return dtype == MMER_BF16 && d == AL_D && Tn + 1 > 32 && Tn + 1 <= 384; }

int mha_fwd_long(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, DropCfg dc, cudaStream_t st,
                 float* lse) {
  LongGeom g;
  long_geom(B, Tn, H, &g);
  CUtensorMap m64, mtail;
  MMER_TRY(long_maps(qkv, 3 * g.F, g, &m64, &mtail));
  const size_t smem = long_smem(g, 3, false);
  MMER_CHECK_ARG(smem <= 232448, "mha_fwd(long): %lld bytes of shared memory needed", (long long)smem);
  auto launch = [&](auto kern, size_t* configured) -> int {
    if (smem > *configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha_fwd_long)");
      *configured = smem;
    }
    // the statistics buffer: (max, 1 / sum) per row, then one keep word per (row, 32-key block) when dropout is on
    uint32_t* keep_g = lse != nullptr ? reinterpret_cast<uint32_t*>(lse + (size_t)B * H * g.S * 2) : nullptr;
    kern<<<(unsigned)((long long)B * H), AL_WARPS * 32, smem, st>>>(m64, mtail, mask, (bf16*)out, probs, reinterpret_cast<float2*>(lse), keep_g, g, dc);
    return 0;
  };
  static size_t conf[4] = {0, 0, 0, 0};
  const bool drop = dc.thr != 0, pr = probs != nullptr;
  if (drop && pr) MMER_TRY(launch(mha_fwd_long_kernel<true, true>, &conf[0]));
  else if (drop) MMER_TRY(launch(mha_fwd_long_kernel<true, false>, &conf[1]));
  else if (pr) MMER_TRY(launch(mha_fwd_long_kernel<false, true>, &conf[2]));
  else MMER_TRY(launch(mha_fwd_long_kernel<false, false>, &conf[3]));
  MMER_LAUNCH_CHECK("mha_fwd_long_kernel");
  return 0;
}

int mha_bwd_long(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, int B, int Tn, int H, DropCfg dc,
                 cudaStream_t st, const float* lse, const void* fwd_out) {
  if (fwd_out == nullptr) lse = nullptr;   // the stored statistics are only useful together with the forward output
  LongGeom g;
  long_geom(B, Tn, H, &g);
  CUtensorMap m64, mtail, d64, dtail;
  MMER_TRY(long_maps(qkv, 3 * g.F, g, &m64, &mtail));
  MMER_TRY(long_maps(dout, g.F, g, &d64, &dtail));
  size_t smem = long_smem(g, 4, true);
  MMER_CHECK_ARG(smem <= 232448, "mha_bwd(long): %lld bytes of shared memory needed", (long long)smem);
  const size_t bits_bytes = (size_t)(g.SR + 32) * ((g.S + 31) / 32) * 4;      // dropout keep-bit map, when it fits
  const bool bits = dc.thr != 0 && smem + bits_bytes <= 232448;
  if (bits) smem += bits_bytes;
  // with the forward's statistics comes its keep-bit map (same buffer, behind the (max, 1 / sum) pairs)
  const uint32_t* keep_g = (lse != nullptr && bits) ? reinterpret_cast<const uint32_t*>(lse + (size_t)B * H * g.S * 2) : nullptr;
  // fewest rounds over the 16-row tiles with 8 or 9 warps
  const int ntile = (g.S + 15) / 16;
  const int nw = ((ntile + 8) / 9 < (ntile + 7) / 8) ? 9 : 8;
  auto launch = [&](auto kern, size_t* configured) -> int {
    if (smem > *configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha_bwd_long)");
      *configured = smem;
    }
    kern<<<(unsigned)((long long)B * H), nw * 32, smem, st>>>(m64, mtail, d64, dtail, mask, (bf16*)dqkv, reinterpret_cast<const float2*>(lse),
                                                            (const bf16*)fwd_out, keep_g, g, dc);
    return 0;
  };
  static size_t conf[4] = {0, 0, 0, 0};
  if (keep_g != nullptr) MMER_TRY(launch(mha_bwd_long_kernel<true, true, true>, &conf[0]));
  else if (bits) MMER_TRY(launch(mha_bwd_long_kernel<true, true, false>, &conf[1]));
  else if (dc.thr != 0) MMER_TRY(launch(mha_bwd_long_kernel<true, false, false>, &conf[2]));
  else MMER_TRY(launch(mha_bwd_long_kernel<false, false, false>, &conf[3]));
  MMER_LAUNCH_CHECK("mha_bwd_long_kernel");
  return 0;
}

}  // namespace mmer
