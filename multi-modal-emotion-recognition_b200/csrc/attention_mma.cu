// Multi-head self-attention over the short token sequence [T video tokens ; audio token], bf16, S = T+1 <= 32.
// Replaces the SDPA inside nn.MultiheadAttention as configured at train2.py:111-118 / train.py:54-57
// (called at train2.py:173-176, train.py:91) and its autograd backward.
//
// One CTA per sample, one warp per head.  The sample's packed in_proj rows ([S][3F] bf16, 3 KB per row at
// F = 512) arrive in shared memory through 1-D bulk (TMA) copies, one per token row, into rows padded by 16 B so
// that ldmatrix is bank-conflict free; completion is counted on an mbarrier.  Per head everything is done with
// warp-level tensor-core MMAs (m16n8k16, bf16 in, fp32 accumulate) on ldmatrix fragments:
//   forward   S = Q K^T, masked softmax in the accumulator fragments (quad shuffles), dropout, O = P V
//   backward  recompute P, dP = dO V^T, dS = P o (dP - rowsum(dP o P)) / sqrt(d), dV = Pd^T dO, dQ = dS K,
//             dK = dS^T Q; the transposed operands (Pd^T, dS^T) are built in registers with movmatrix
// Results overwrite operand slots that are dead by then (O -> Q slot; dV -> V slot, dK -> K slot, dQ -> dO slot),
// so whole token rows leave through bulk shared->global copies.  HBM traffic is the algorithmic minimum: every
// input byte is read once, every output byte written once, all as >= 1 KB contiguous bursts.
// Rows/keys beyond S are handled by clamping fragment addresses to row S-1 (finite data) and zeroing their
// probabilities, so no shared memory beyond the S real rows is needed.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

extern int g_debug[16];
static constexpr int MMA_WARPS = 8;

struct MmaGeom {
  int S, F, Tn, H;
  uint32_t in_row, in_stride;   // bytes of one packed qkv row, padded smem stride
  uint32_t do_row, do_stride;   // bytes of one dO / out row, padded smem stride
};

// Byte offsets inside a [rows][D] bf16 head tile in shared memory, as the fragment loads and stores of this file need
// them.  Every access pattern is "a row that depends on the lane (clamped to S - 1 for loads: rows beyond S hold no
// data) and a 16-byte chunk = a compile-time step + a lane term":
//   a_off(i, ks)   ldmatrix.x4 of row tile i (16 rows), 16-column step ks: A operands, and the transposed B operands
//   b_off(nt, k2)  ldmatrix.x4 of key tile nt (8 rows), 32-column step k2: B operands (K, V)
//   st_base(tile, rowbase) / st_addr(base, nd)   the lane's bf16 pair of row rowbase + g, 8-column block nd (swizzled tiles)
struct PadAddr {      // row-padded packed rows (bulk-copied whole token rows): offset = row * stride + col * 2
  static constexpr bool kStaticRows = false;   // a tile has exactly S rows: stores are guarded per row
  uint32_t stride;
  int lane, S;
  __device__ __forceinline__ PadAddr(uint32_t stride_, int lane_, int S_) : stride(stride_), lane(lane_), S(S_) {}
  __device__ __forceinline__ uint32_t off(int row, int col) const { return (uint32_t)row * stride + (uint32_t)col * 2u; }
  __device__ __forceinline__ uint32_t a_off(int i, int ks) const {
    return off(min(i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S - 1), ks * 16 + (lane >> 4) * 8);
  }
  __device__ __forceinline__ uint32_t b_off(int nt, int k2) const {
    return off(min(nt * 8 + (lane & 7), S - 1), k2 * 32 + (lane >> 3) * 8);
  }
};
// One 128-byte row per token, 128B-swizzled as TMA writes it (D = 64, tiles 1024-byte aligned and padded to whole
// 8-row groups): conflict-free ldmatrix.  offset = row * 128 + ((chunk ^ (row & 7)) << 4) + byte in chunk.  The lane
// terms are computed ONCE per kernel (they depend on the lane and S only, not on the head): a chunk is a compile-time
// step XOR a lane term, so every address in the per-head code is `precomputed ^ immediate` (+ the tile base).
struct SwzAddr {
  static constexpr bool kStaticRows = true;    // rows up to the next multiple of 8 exist in the tile (never stored to HBM)
  uint32_t pa[2], pb[4], pst;
  __device__ __forceinline__ SwzAddr(int lane, int S) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t row = (uint32_t)min(i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S - 1);
      pa[i] = row * 128u + ((((uint32_t)lane >> 4) ^ (row & 7u)) << 4);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t row = (uint32_t)min(nt * 8 + (lane & 7), S - 1);
      pb[nt] = row * 128u + ((((uint32_t)lane >> 3) ^ (row & 7u)) << 4);
    }
    const uint32_t g = (uint32_t)lane >> 2, t = (uint32_t)lane & 3u;
    pst = g * 128u + (g << 4) + t * 4u;
  }
  __device__ __forceinline__ uint32_t off(int row, int col) const {
    return (uint32_t)row * 128u + (((((uint32_t)col >> 3) ^ (uint32_t)row) & 7u) << 4) + ((uint32_t)col & 7u) * 2u;
  }
  __device__ __forceinline__ uint32_t a_off(int i, int ks) const { return pa[i] ^ ((uint32_t)ks << 5); }
  __device__ __forceinline__ uint32_t b_off(int nt, int k2) const { return pb[nt] ^ ((uint32_t)k2 << 6); }
  __device__ __forceinline__ uint32_t st_base(uint32_t tile, int rowbase) const { return tile + (uint32_t)rowbase * 128u + pst; }
  __device__ __forceinline__ uint32_t st_addr(uint32_t base, int nd) const { return base ^ ((uint32_t)nd << 4); }
};
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

template <int NT>
__device__ __forceinline__ uint32_t key_valid_bits(const uint8_t* __restrict__ mask, int b, int Tn, int S, int t) {
  uint32_t bits = 0;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + t * 2 + e;
      bool ok = j < S;
      if (ok && j < Tn && mask != nullptr) ok = mask[(long long)b * Tn + j] == 0;
      bits |= (ok ? 1u : 0u) << (nt * 2 + e);
    }
  return bits;
}

// scores + masked softmax for one head: p[mt][nt][..] = softmax_j(q_i . k_j / sqrt(D)), fragment layout of the
// m16n8 accumulators (row g / g+8, columns nt*8 + t*2 + {0,1}); invalid keys get exactly 0.
template <int D, int MT, int NT, class AD>
__device__ __forceinline__ void scores_softmax(uint32_t qbase, uint32_t kbase, AD ad, int S, int lane,
                                               uint32_t kvalid, float (&p)[MT][NT][4]) {
  constexpr int KS = D / 16;
  uint32_t kf[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int k2 = 0; k2 < KS / 2; ++k2)
      ldsm_x4(kbase + ad.b_off(nt, k2), kf[nt][2 * k2][0], kf[nt][2 * k2][1], kf[nt][2 * k2 + 1][0], kf[nt][2 * k2 + 1][1]);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) p[mt][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(qbase + ad.a_off(mt, ks), a0, a1, a2, a3);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(p[mt][nt], a0, a1, a2, a3, kf[nt][ks][0], kf[nt][ks][1]);
    }
  }
  const float sl2 = rsqrtf((float)D) * 1.4426950408889634f;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if ((kvalid >> (nt * 2 + e)) & 1u) m = fmaxf(m, p[mt][nt][r * 2 + e]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float x = ((kvalid >> (nt * 2 + e)) & 1u) ? ex2_approx((p[mt][nt][r * 2 + e] - m) * sl2) : 0.f;
          p[mt][nt][r * 2 + e] = x;
          sum += x;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) p[mt][nt][r * 2 + e] *= inv;
    }
}

// A fragment (16 x 16, rows m, columns k) of X^T from the A fragment of X covering the same 16 x 16 block
__device__ __forceinline__ void trans_frag(const uint32_t (&x)[4], uint32_t (&y)[4]) {
  y[0] = movmatrix_t(x[0]);
  y[1] = movmatrix_t(x[2]);
  y[2] = movmatrix_t(x[1]);
  y[3] = movmatrix_t(x[3]);
}

// accumulator fragments [MT][NT][4] of a (queries x keys) matrix -> bf16 A fragments over 16-key steps
template <int MT, int NT>
__device__ __forceinline__ void pack_rows(const float (&c)[MT][NT][4], uint32_t (&a)[MT][(NT + 1) / 2][4]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < (NT + 1) / 2; ++ks) {
      a[mt][ks][0] = pack_bf16x2(c[mt][2 * ks][0], c[mt][2 * ks][1]);
      a[mt][ks][1] = pack_bf16x2(c[mt][2 * ks][2], c[mt][2 * ks][3]);
      if (2 * ks + 1 < NT) {
        a[mt][ks][2] = pack_bf16x2(c[mt][2 * ks + 1][0], c[mt][2 * ks + 1][1]);
        a[mt][ks][3] = pack_bf16x2(c[mt][2 * ks + 1][2], c[mt][2 * ks + 1][3]);
      } else {
        a[mt][ks][2] = 0u;
        a[mt][ks][3] = 0u;
      }
    }
}

// acc[D/8][4] (+)= A-fragments(a, 16 x 16*KSTEPS) . X[rows 16*ks.. , D columns] with X row-major in smem (ldmatrix.trans)
template <int D, int KSTEPS, class AD>
__device__ __forceinline__ void mma_rows_x(float (&acc)[D / 8][4], const uint32_t (&a)[KSTEPS][4], uint32_t xbase,
                                           AD ad, int S, int lane) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
    for (int n2 = 0; n2 < D / 16; ++n2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(xbase + ad.a_off(ks, n2), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * n2], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
      mma_bf16_16816(acc[2 * n2 + 1], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
    }
}

// store a 16 x D accumulator tile as bf16 rows (row0 + g, row0 + g + 8) of the shared-memory tile at address `tile`.
// Rows that do not exist are skipped: beyond S in the padded-row layout (per lane), beyond the tile's NT * 8 rows in the
// swizzled one (decided at compile time; rows S .. NT * 8 - 1 are padding that no TMA store reads).
template <int D, int NT, class AD>
__device__ __forceinline__ void store_tile(uint32_t tile, const AD& ad, int row0, int lane, const float (&acc)[D / 8][4]) {
  if constexpr (AD::kStaticRows) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int rowbase = row0 + 8 * r;
      if (rowbase < NT * 8) {
        const uint32_t base = ad.st_base(tile, rowbase);
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) sts_b32(ad.st_addr(base, nd), pack_bf16x2(acc[nd][2 * r], acc[nd][2 * r + 1]));
      }
    }
  } else {
    uint8_t* base = reinterpret_cast<uint8_t*>(__cvta_shared_to_generic(tile));
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row0 + g + 8 * r;
      if (row < ad.S) {
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd)
          *reinterpret_cast<uint32_t*>(base + ad.off(row, nd * 8 + t * 2)) = pack_bf16x2(acc[nd][2 * r], acc[nd][2 * r + 1]);
      }
    }
  }
}

// One head of the forward pass on fragments: S = Q K^T, masked softmax, dropout, O = P V.  q/k/v are shared-memory
// addresses of [S][D] tiles laid out per the address policy `ad`; O (bf16) is written to o_ptr with the same policy.
template <int D, int MT, int NT, class AD>
__device__ __forceinline__ void mha_fwd_head(uint32_t qbase, uint32_t kbase, uint32_t vbase, uint32_t o_a, const AD& ad,
                                             int S, int lane, uint32_t kvalid, long long bh, float* __restrict__ probs,
                                             DropCfg dc) {
  const int g = lane >> 2, t = lane & 3;
  float p[MT][NT][4];
  scores_softmax<D, MT, NT>(qbase, kbase, ad, S, lane, kvalid, p);
  if (probs != nullptr || dc.thr) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = mt * 16 + g + 8 * r;
        if (i < S) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int j = nt * 8 + t * 2;
            if (probs != nullptr) {
              if (j < S) probs[(bh * S + i) * S + j] = p[mt][nt][r * 2];
              if (j + 1 < S) probs[(bh * S + i) * S + j + 1] = p[mt][nt][r * 2 + 1];
            }
            if (dc.thr) {
              float f0, f1;
              // element index (bh*S+i) * NT*8 + j.  The hoisted form is 2 % faster in the TMA-tile kernel and 20 % slower
              // in the bulk-row one (measured; register allocation), hence the switch on the tile addressing.
              if constexpr (std::is_same<AD, SwzAddr>::value) drop2_at(dc, (uint32_t)((bh * S + i) * NT + nt), drop_mult(t), f0, f1);
              else drop2(dc, att_drop_index(bh * S + i, j, NT * 8), f0, f1);
              p[mt][nt][r * 2] *= f0;
              p[mt][nt][r * 2 + 1] *= f1;
            }
          }
        }
      }
  }
  uint32_t pa[MT][(NT + 1) / 2][4];
  pack_rows<MT, NT>(p, pa);
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float o[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[nd][i] = 0.f;
    mma_rows_x<D, (NT + 1) / 2>(o, pa[mt], vbase, ad, S, lane);
    store_tile<D, NT>(o_a, ad, mt * 16, lane, o);   // O_h overwrites the dead Q_h slot
  }
}

template <int D, int MT, int NT>
__global__ void __launch_bounds__(MMA_WARPS * 32)
mha_fwd_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ mask, bf16* __restrict__ out,
                   float* __restrict__ probs, MmaGeom gm, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int S = gm.S, F = gm.F, H = gm.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)S * gm.in_stride + 15) & ~size_t(15)));
  const uint32_t bar_a = smem_u32(bar);
  const uint32_t in_a = smem_u32(smem);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mbar_expect_tx(bar_a, (uint32_t)S * gm.in_row);
    __syncwarp();
    for (int r = lane; r < S; r += 32)
      bulk_g2s(in_a + r * gm.in_stride, qkv + ((long long)b * S + r) * 3 * F, gm.in_row, bar_a);
  }
  const int g = lane >> 2, t = lane & 3;
  const uint32_t kvalid = key_valid_bits<NT>(mask, b, gm.Tn, S, t);
  mbar_wait(bar_a, 0);

  for (int h = warp; h < H; h += MMA_WARPS) {
    const uint32_t qbase = in_a + h * D * 2, kbase = qbase + F * 2, vbase = kbase + F * 2;
    mha_fwd_head<D, MT, NT>(qbase, kbase, vbase, in_a + h * D * 2, PadAddr(gm.in_stride, lane, S), S, lane, kvalid,
                            (long long)b * H + h, probs, dc);
  }
  fence_async_smem();
  __syncthreads();
  if (warp == 0) {
    for (int r = lane; r < S; r += 32) bulk_s2g(out + ((long long)b * S + r) * F, in_a + r * gm.in_stride, gm.do_row);
    bulk_commit();
    bulk_wait_read0();
  }
}

// One head of the backward pass on fragments.  q/k/v tiles follow policy `ain`, the dO tile policy `ado`.
// Outputs overwrite dead operand tiles: dV -> dv_ptr (ain), dK -> dk_ptr (ain), dQ -> dq_ptr (ado).
template <int D, int MT, int NT, class AIN, class ADO>
__device__ __forceinline__ void mha_bwd_head(uint32_t qbase, uint32_t kbase, uint32_t vbase, uint32_t dobase,
                                             uint32_t dq_a, uint32_t dk_a, uint32_t dv_a, const AIN& ain,
                                             const ADO& ado, int S, int lane, uint32_t kvalid, long long bh, DropCfg dc) {
  constexpr int KS = D / 16;
  constexpr int MTK = (NT + 1) / 2;   // 16-row tiles over keys
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf((float)D);
  float p[MT][NT][4];
  scores_softmax<D, MT, NT>(qbase, kbase, ain, S, lane, kvalid, p);
  // dP = dO V^T
  float dp[MT][NT][4];
  {
    uint32_t vf[NT][KS][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int k2 = 0; k2 < KS / 2; ++k2)
        ldsm_x4(vbase + ain.b_off(nt, k2), vf[nt][2 * k2][0], vf[nt][2 * k2][1], vf[nt][2 * k2 + 1][0], vf[nt][2 * k2 + 1][1]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) dp[mt][nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(dobase + ado.a_off(mt, ks), a0, a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(dp[mt][nt], a0, a1, a2, a3, vf[nt][ks][0], vf[nt][ks][1]);
      }
    }
  }
  // p <- Pd = P o dropout (what multiplied V in the forward pass); dp <- dS.  Query rows >= S are zeroed: they
  // are reduction indices of dV and dK.
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = mt * 16 + g + 8 * r;
      const bool row_ok = i < S;
      float f[NT][2];
      float dot = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        f[nt][0] = 1.f;
        f[nt][1] = 1.f;
        if (dc.thr && row_ok) drop2_at(dc, (uint32_t)((bh * S + i) * NT + nt), drop_mult(t), f[nt][0], f[nt][1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float dpm = dp[mt][nt][r * 2 + e] * f[nt][e];
          dp[mt][nt][r * 2 + e] = dpm;
          dot = fmaf(dpm, p[mt][nt][r * 2 + e], dot);
        }
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pv = p[mt][nt][r * 2 + e];
          dp[mt][nt][r * 2 + e] = row_ok ? pv * (dp[mt][nt][r * 2 + e] - dot) * scale : 0.f;
          p[mt][nt][r * 2 + e] = row_ok ? pv * f[nt][e] : 0.f;
        }
    }
  // ---- dV = Pd^T dO.  The A fragments of a transposed operand are the 8x8 blocks of the original's fragments,
  // each transposed in registers (movmatrix) and with the two off-diagonal blocks swapped: no shared memory.
  uint32_t pda[MT][(NT + 1) / 2][4];
  pack_rows<MT, NT>(p, pda);
#pragma unroll
  for (int mk = 0; mk < MTK; ++mk) {
    uint32_t a[MT][4];
#pragma unroll
    for (int kq = 0; kq < MT; ++kq) trans_frag(pda[kq][mk], a[kq]);
    float acc[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nd][i] = 0.f;
    mma_rows_x<D, MT>(acc, a, dobase, ado, S, lane);
    store_tile<D, NT>(dv_a, ain, mk * 16, lane, acc);   // dV_h -> dead V_h slot
  }
  // ---- dS as A fragments (for dQ); its transpose is built the same way (for dK)
  uint32_t dsa[MT][(NT + 1) / 2][4];
  pack_rows<MT, NT>(dp, dsa);
  // ---- dQ = dS K -> dead dO_h slot
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float acc[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nd][i] = 0.f;
    mma_rows_x<D, (NT + 1) / 2>(acc, dsa[mt], kbase, ain, S, lane);
    store_tile<D, NT>(dq_a, ado, mt * 16, lane, acc);
  }
  // ---- dK = dS^T Q -> dead K_h slot, one 16-key tile at a time (the accumulators of one tile are live, not of all).
  // Every lane has finished reading K_h (dQ above) before the first tile overwrites it; the later tiles read Q only.
  __syncwarp();
#pragma unroll
  for (int mk = 0; mk < MTK; ++mk) {
    uint32_t a[MT][4];
#pragma unroll
    for (int kq = 0; kq < MT; ++kq) trans_frag(dsa[kq][mk], a[kq]);
    float acc[D / 8][4];
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nd][i] = 0.f;
    mma_rows_x<D, MT>(acc, a, qbase, ain, S, lane);
    store_tile<D, NT>(dk_a, ain, mk * 16, lane, acc);
  }
  __syncwarp();
}

template <int D, int MT, int NT>
__global__ void __launch_bounds__(MMA_WARPS * 32, 2)
mha_bwd_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ mask, const bf16* __restrict__ dout,
                   bf16* __restrict__ dqkv, float* __restrict__ dbias, int B, MmaGeom gm, DropCfg dc) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int KS = D / 16;
  constexpr int MTK = (NT + 1) / 2;   // 16-row tiles over keys
  const int S = gm.S, F = gm.F, H = gm.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t in_bytes = ((size_t)S * gm.in_stride + 15) & ~size_t(15);
  const size_t do_bytes = ((size_t)S * gm.do_stride + 15) & ~size_t(15);
  uint8_t* do_s = smem + in_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(do_s + do_bytes);
  float* colacc = reinterpret_cast<float*>(bar + 2);   // [3F] running column sums of dqkv (in_proj bias gradient)
  const uint32_t bar_a = smem_u32(bar), in_a = smem_u32(smem), do_a = smem_u32(do_s);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  if (dbias != nullptr)
    for (int c = threadIdx.x; c < 3 * F; c += blockDim.x) colacc[c] = 0.f;
  __syncthreads();
  // persistent over samples: the loads of a sample are issued as soon as the previous sample's stores have drained
  // the shared-memory rows (a second CTA on the SM covers the gap)
  auto issue_loads = [&](int b) {
    if (lane == 0) mbar_expect_tx(bar_a, (uint32_t)S * (gm.in_row + gm.do_row));
    __syncwarp();
    for (int r = lane; r < S; r += 32) {
      bulk_g2s(in_a + r * gm.in_stride, qkv + ((long long)b * S + r) * 3 * F, gm.in_row, bar_a);
      bulk_g2s(do_a + r * gm.do_stride, dout + ((long long)b * S + r) * F, gm.do_row, bar_a);
    }
  };
  if (warp == 0 && (int)blockIdx.x < B) issue_loads(blockIdx.x);
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf((float)D);
  uint32_t phase = 0;
#ifdef MMER_ATT_PROFILE
  long long pt[4] = {0, 0, 0, 0}, pc = clock64();
#define ATT_TICK(i) do { long long _n = clock64(); pt[i] += _n - pc; pc = _n; } while (0)
#else
#define ATT_TICK(i)
#endif
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
  const uint32_t kvalid = key_valid_bits<NT>(mask, b, gm.Tn, S, t);
  ATT_TICK(3);
  mbar_wait(bar_a, phase);
  ATT_TICK(0);
  phase ^= 1;

  for (int h = warp; h < H; h += MMA_WARPS) {
    const uint32_t qbase = in_a + h * D * 2, kbase = qbase + F * 2, vbase = kbase + F * 2, dobase = do_a + h * D * 2;
    mha_bwd_head<D, MT, NT>(qbase, kbase, vbase, dobase, dobase, kbase, vbase, PadAddr(gm.in_stride, lane, S),
                            PadAddr(gm.do_stride, lane, S), S, lane, kvalid, (long long)b * H + h, dc);
  }
  ATT_TICK(1);
  fence_async_smem();
  __syncthreads();
  ATT_TICK(2);
  if (warp == 0) {
    for (int r = lane; r < S; r += 32) {
      bf16* drow = dqkv + ((long long)b * S + r) * 3 * F;
      bulk_s2g(drow, do_a + r * gm.do_stride, gm.do_row);                            // dQ
      bulk_s2g(drow + F, in_a + r * gm.in_stride + gm.do_row, 2 * gm.do_row);        // dK, dV
    }
    bulk_commit();
  }
  if (dbias != nullptr) {
    // column sums of this sample's dQ | dK | dV rows (what is stored), two columns per thread
    for (int p2 = threadIdx.x; p2 < 3 * F / 2; p2 += blockDim.x) {
      const int c = p2 * 2;
      const uint8_t* src = c < F ? do_s + c * 2 : smem + gm.do_row + (c - F) * 2;
      const uint32_t stride = c < F ? gm.do_stride : gm.in_stride;
      float a0 = 0.f, a1 = 0.f;
      for (int r = 0; r < S; ++r) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + (size_t)r * stride));
        a0 += v.x;
        a1 += v.y;
      }
      colacc[c] += a0;
      colacc[c + 1] += a1;
    }
  }
  if (warp == 0) bulk_wait_read0();
  __syncthreads();   // rows drained by the stores and read by the column sums: the next sample may land
  if (warp == 0 && b + (int)gridDim.x < B) issue_loads(b + gridDim.x);
  }  // sample loop
#ifdef MMER_ATT_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("mha_bwd warp %d: wait_load %lld compute %lld barrier %lld store+colsum+issue %lld\n", warp, pt[0], pt[1], pt[2], pt[3]);
#endif
  if (dbias != nullptr) {
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * F; c += blockDim.x) atomicAdd(dbias + c, colacc[c]);
  }
}

// =========================================================================================================
// Warp-pipelined variant for head size 64.  The unit of work is one (sample, head): its Q, K, V (and dO) tiles are
// [S][64] bf16 = S rows of 128 B, fetched by ONE 2-D TMA box each straight from the packed activation matrices into
// 128B-swizzled shared-memory tiles (conflict-free ldmatrix without padding), and results leave by TMA box stores.
// Every warp runs its own load -> compute -> store loop on its own mbarrier, so an SM has as many independent
// pipelines in flight as it has warps (16-24) instead of one per CTA: loads of some units overlap the math and the
// stores of others.  A warp keeps the same head for all its units (grid stride is a multiple of H), so the in_proj
// bias gradient accumulates in registers.
// =========================================================================================================
static constexpr int TMA_FWD_WARPS = 8;
static constexpr int TMA_BWD_WARPS = 8;

template <int MT, int NT>
__global__ void __launch_bounds__(TMA_FWD_WARPS * 32, 3)
mha_fwd_tma_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                   const uint8_t* __restrict__ mask, float* __restrict__ probs, int B, int Tn, int H, uint32_t tile_bytes,
                   DropCfg dc) {
  constexpr int D = 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int S = Tn + 1, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem + (size_t)warp * 3 * tile_bytes;            // Q | K | V tiles of this warp
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)TMA_FWD_WARPS * 3 * tile_bytes) + warp;
  const uint32_t bar_a = smem_u32(bar), q_a = smem_u32(my), k_a = q_a + tile_bytes, v_a = k_a + tile_bytes;
  pdl_trigger();
  if (lane == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  __syncwarp();
  pdl_wait();
  const int t = lane & 3;
  const long long units = (long long)B * H;
  const long long stride = (long long)gridDim.x * TMA_FWD_WARPS;
  uint32_t phase = 0;
  const SwzAddr ad(lane, S);
  for (long long u = (long long)blockIdx.x * TMA_FWD_WARPS + warp; u < units; u += stride) {
    const int b = (int)(u / H), h = (int)(u % H);
    if (lane == 0) {
      mbar_expect_tx(bar_a, 3u * (uint32_t)S * 128u);
      tma_load_2d(q_a, &tm_qkv, bar_a, h * D, b * S);
      tma_load_2d(k_a, &tm_qkv, bar_a, F + h * D, b * S);
      tma_load_2d(v_a, &tm_qkv, bar_a, 2 * F + h * D, b * S);
    }
    const uint32_t kvalid = key_valid_bits<NT>(mask, b, Tn, S, t);
    mbar_wait(bar_a, phase);
    phase ^= 1;
    mha_fwd_head<D, MT, NT>(q_a, k_a, v_a, q_a, ad, S, lane, kvalid, u, probs, dc);
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tm_out, q_a, h * D, b * S);     // O_h sits in the Q tile
      bulk_commit();
      bulk_wait_read0();                            // the tile is reloaded next iteration
    }
    __syncwarp();
  }
}

template <int MT, int NT>
__global__ void __launch_bounds__(TMA_BWD_WARPS * 32, 2)
mha_bwd_tma_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dqkv, const uint8_t* __restrict__ mask,
                   float* __restrict__ dbias, int B, int Tn, int H, uint32_t tile_bytes, DropCfg dc) {
  constexpr int D = 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int S = Tn + 1, F = H * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem + (size_t)warp * 4 * tile_bytes;            // Q | K | V | dO tiles of this warp
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)TMA_BWD_WARPS * 4 * tile_bytes) + warp;
  const uint32_t bar_a = smem_u32(bar), q_a = smem_u32(my), k_a = q_a + tile_bytes, v_a = k_a + tile_bytes,
                 do_a = v_a + tile_bytes;
  pdl_trigger();
  if (lane == 0) {
    mbar_init(bar_a, 1);
    mbar_init_fence();
  }
  __syncwarp();
  pdl_wait();
  const int t = lane & 3;
  const long long units = (long long)B * H;
  const long long stride = (long long)gridDim.x * TMA_BWD_WARPS;   // host makes it a multiple of H: h is fixed per warp
  float cs[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};           // column sums of dQ | dK | dV, columns 2*lane, 2*lane+1
  int my_h = -1;
  uint32_t phase = 0;
  const SwzAddr ad(lane, S);
  // column sums of the stored tiles: the lane's bf16 pair (columns 2 * lane, 2 * lane + 1) of row r sits in chunk
  // (lane >> 2) ^ (r & 7); rows r, r + 8, r + 16, ... share the swizzle, so the loop is unrolled over r & 7
  const uint32_t pcs = (((uint32_t)lane >> 2) << 4) + ((uint32_t)lane & 3u) * 4u;
  for (long long u = (long long)blockIdx.x * TMA_BWD_WARPS + warp; u < units; u += stride) {
    const int b = (int)(u / H), h = (int)(u % H);
    my_h = h;
    if (lane == 0) {
      mbar_expect_tx(bar_a, 4u * (uint32_t)S * 128u);
      tma_load_2d(q_a, &tm_qkv, bar_a, h * D, b * S);
      tma_load_2d(k_a, &tm_qkv, bar_a, F + h * D, b * S);
      tma_load_2d(v_a, &tm_qkv, bar_a, 2 * F + h * D, b * S);
      tma_load_2d(do_a, &tm_do, bar_a, h * D, b * S);
    }
    const uint32_t kvalid = key_valid_bits<NT>(mask, b, Tn, S, t);
    mbar_wait(bar_a, phase);
    phase ^= 1;
    mha_bwd_head<D, MT, NT>(q_a, k_a, v_a, do_a, do_a, k_a, v_a, ad, ad, S, lane, kvalid, u, dc);
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tm_dqkv, do_a, h * D, b * S);            // dQ_h (dO tile)
      tma_store_2d(&tm_dqkv, k_a, F + h * D, b * S);         // dK_h
      tma_store_2d(&tm_dqkv, v_a, 2 * F + h * D, b * S);     // dV_h
      bulk_commit();
    }
    if (dbias != nullptr) {
      // column sums of what was stored, two columns per lane, while the stores drain
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const uint32_t tile = (m == 0 ? do_a : m == 1 ? k_a : v_a) + pcs;
#pragma unroll
        for (int r7 = 0; r7 < 8; ++r7) {
          const uint32_t a = (tile ^ ((uint32_t)r7 << 4)) + (uint32_t)r7 * 128u;
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            if (k < NT - 1 || r7 + 8 * k < S) {     // only the last 8-row group is partial
              const uint32_t v = lds_b32(a + (uint32_t)k * 1024u);
              cs[m][0] += __uint_as_float(v << 16);
              cs[m][1] += __uint_as_float(v & 0xffff0000u);
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  }
  if (dbias != nullptr && my_h >= 0) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      atomicAdd(dbias + m * F + my_h * D + 2 * lane, cs[m][0]);
      atomicAdd(dbias + m * F + my_h * D + 2 * lane + 1, cs[m][1]);
    }
  }
}

template <typename K>
static int tma_set_smem(K kern, size_t smem, size_t* configured) {
  if (smem > *configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha tma)");
    *configured = smem;
  }
  return 0;
}

template <int MT, int NT>
static int fwd_tma_launch(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, DropCfg dc,
                          cudaStream_t st) {
  const int S = Tn + 1, F = H * 64;
  const uint32_t tile_bytes = (uint32_t)((S + 7) / 8) * 1024u;
  CUtensorMap tq, to;
  MMER_TRY(make_tma_map_bf16(&tq, qkv, (uint64_t)3 * F, (uint64_t)B * S, (uint64_t)3 * F, 64, (uint32_t)S));
  MMER_TRY(make_tma_map_bf16(&to, out, (uint64_t)F, (uint64_t)B * S, (uint64_t)F, 64, (uint32_t)S));
  const size_t smem = (size_t)TMA_FWD_WARPS * 3 * tile_bytes + TMA_FWD_WARPS * 8;
  static size_t configured = 0;
  static int bps = 0;
  auto kern = mha_fwd_tma_kernel<MT, NT>;
  if (smem > configured) bps = 0;
  MMER_TRY(tma_set_smem(kern, smem, &configured));
  if (bps == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TMA_FWD_WARPS * 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy(mha_fwd_tma)");
    if (bps < 1) bps = 1;
  }
  const long long units = (long long)B * H;
  long long grid = (units + TMA_FWD_WARPS - 1) / TMA_FWD_WARPS;
  const long long cap = (long long)sm_count() * bps;
  if (grid > cap) grid = cap;
  cudaError_t le = launch_dep(kern, dim3((unsigned)grid), dim3(TMA_FWD_WARPS * 32), smem, st, 1, tq, to, mask, probs, B, Tn, H,
                              tile_bytes, dc);
  if (le != cudaSuccess) return cuda_fail(le, "launch(mha_fwd_tma)");
  MMER_LAUNCH_CHECK("mha_fwd_tma_kernel");
  return 0;
}
template <int MT, int NT>
static int bwd_tma_launch(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B, int Tn,
                          int H, DropCfg dc, cudaStream_t st) {
  const int S = Tn + 1, F = H * 64;
  const uint32_t tile_bytes = (uint32_t)((S + 7) / 8) * 1024u;
  CUtensorMap tq, td, tg;
  MMER_TRY(make_tma_map_bf16(&tq, qkv, (uint64_t)3 * F, (uint64_t)B * S, (uint64_t)3 * F, 64, (uint32_t)S));
  MMER_TRY(make_tma_map_bf16(&td, dout, (uint64_t)F, (uint64_t)B * S, (uint64_t)F, 64, (uint32_t)S));
  MMER_TRY(make_tma_map_bf16(&tg, dqkv, (uint64_t)3 * F, (uint64_t)B * S, (uint64_t)3 * F, 64, (uint32_t)S));
  const size_t smem = (size_t)TMA_BWD_WARPS * 4 * tile_bytes + TMA_BWD_WARPS * 8;
  static size_t configured = 0;
  static int bps = 0;
  auto kern = mha_bwd_tma_kernel<MT, NT>;
  if (smem > configured) bps = 0;
  MMER_TRY(tma_set_smem(kern, smem, &configured));
  if (bps == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TMA_BWD_WARPS * 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy(mha_bwd_tma)");
    if (bps < 1) bps = 1;
  }
  // total warps must be a multiple of H so that every warp keeps one head (register column sums)
  const long long units = (long long)B * H;
  long long grid = (units + TMA_BWD_WARPS - 1) / TMA_BWD_WARPS;
  const long long cap = (long long)sm_count() * bps;
  if (grid > cap) grid = cap;
  if (dbias != nullptr) {
    while (grid > 1 && (grid * TMA_BWD_WARPS) % H != 0) --grid;
    MMER_CHECK_ARG((grid * TMA_BWD_WARPS) % H == 0 || units <= grid * TMA_BWD_WARPS,
                   "mha_bwd: cannot tile %d heads over %d-warp CTAs for the fused bias gradient", H, TMA_BWD_WARPS);
  }
  cudaError_t le = launch_dep(kern, dim3((unsigned)grid), dim3(TMA_BWD_WARPS * 32), smem, st, 1, tq, td, tg, mask, dbias, B, Tn, H,
                              tile_bytes, dc);
  if (le != cudaSuccess) return cuda_fail(le, "launch(mha_bwd_tma)");
  MMER_LAUNCH_CHECK("mha_bwd_tma_kernel");
  return 0;
}

static MmaGeom make_geom(int Tn, int H, int D) {
  MmaGeom g;
  g.Tn = Tn; g.S = Tn + 1; g.H = H; g.F = H * D;
  g.in_row = (uint32_t)(3 * g.F * 2); g.in_stride = g.in_row + 16;
  g.do_row = (uint32_t)(g.F * 2); g.do_stride = g.do_row + 16;
  return g;
}
static size_t fwd_smem(const MmaGeom& g) { return (((size_t)g.S * g.in_stride + 15) & ~size_t(15)) + 16; }
static size_t bwd_smem(const MmaGeom& g) {
  return (((size_t)g.S * g.in_stride + 15) & ~size_t(15)) + (((size_t)g.S * g.do_stride + 15) & ~size_t(15)) +
         16 + (size_t)3 * g.F * sizeof(float);
}

template <typename K>
static int set_smem(K kern, size_t smem, size_t* configured) {
  if (smem > *configured) {
    MMER_CHECK_ARG(smem <= 232448, "mha(mma): %lld bytes of shared memory needed, over the 227 KB limit", (long long)smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mha mma)");
    *configured = smem;
  }
  return 0;
}

template <int D, int MT, int NT>
static int fwd_launch(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, const MmaGeom& g, DropCfg dc,
                      cudaStream_t st) {
  static size_t configured = 0;
  auto kern = mha_fwd_mma_kernel<D, MT, NT>;
  const size_t smem = fwd_smem(g);
  MMER_TRY(set_smem(kern, smem, &configured));
  kern<<<B, MMA_WARPS * 32, smem, st>>>((const bf16*)qkv, mask, (bf16*)out, probs, g, dc);
  MMER_LAUNCH_CHECK("mha_fwd_mma_kernel");
  return 0;
}
template <int D, int MT, int NT>
static int bwd_launch(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B,
                      const MmaGeom& g, DropCfg dc, cudaStream_t st) {
  static size_t configured = 0;
  static int bps = 0;
  auto kern = mha_bwd_mma_kernel<D, MT, NT>;
  const size_t smem = bwd_smem(g);
  if (smem > configured) bps = 0;
  MMER_TRY(set_smem(kern, smem, &configured));
  if (bps == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, MMA_WARPS * 32, smem);
    if (e != cudaSuccess) return cuda_fail(e, "occupancy(mha_bwd_mma)");
    if (bps < 1) bps = 1;
  }
  const long long cap = (long long)sm_count() * bps;
  const int grid = (int)(B < cap ? B : cap);
  kern<<<grid, MMA_WARPS * 32, smem, st>>>((const bf16*)qkv, mask, (const bf16*)dout, (bf16*)dqkv, dbias, B, g, dc);
  MMER_LAUNCH_CHECK("mha_bwd_mma_kernel");
  return 0;
}

template <int D>
static int fwd_d(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, const MmaGeom& g, DropCfg dc,
                 cudaStream_t st) {
  if (g.S <= 8) return fwd_launch<D, 1, 1>(qkv, mask, out, probs, B, g, dc, st);
  if (g.S <= 16) return fwd_launch<D, 1, 2>(qkv, mask, out, probs, B, g, dc, st);
  if (g.S <= 24) return fwd_launch<D, 2, 3>(qkv, mask, out, probs, B, g, dc, st);
  return fwd_launch<D, 2, 4>(qkv, mask, out, probs, B, g, dc, st);
}
template <int D>
static int bwd_d(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B,
                 const MmaGeom& g, DropCfg dc, cudaStream_t st) {
  if (g.S <= 8) return bwd_launch<D, 1, 1>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
  if (g.S <= 16) return bwd_launch<D, 1, 2>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
  if (g.S <= 24) return bwd_launch<D, 2, 3>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
  return bwd_launch<D, 2, 4>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
}

// bf16, S = Tn + 1 <= 32, d in {32, 64}
int mha_fwd_mma(const void* qkv, const uint8_t* mask, void* out, float* probs, int B, int Tn, int H, int d, DropCfg dc,
                cudaStream_t st) {
  const MmaGeom g = make_geom(Tn, H, d);
  MMER_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "mha_fwd: qkv/out must be 16-byte aligned");
  if (d == 64 && !g_debug[MMER_DEBUG_ATT_ROWS]) {   // warp-pipelined TMA-tile kernels
    if (g.S <= 8) return fwd_tma_launch<1, 1>(qkv, mask, out, probs, B, Tn, H, dc, st);
    if (g.S <= 16) return fwd_tma_launch<1, 2>(qkv, mask, out, probs, B, Tn, H, dc, st);
    if (g.S <= 24) return fwd_tma_launch<2, 3>(qkv, mask, out, probs, B, Tn, H, dc, st);
    return fwd_tma_launch<2, 4>(qkv, mask, out, probs, B, Tn, H, dc, st);
  }
  return d == 64 ? fwd_d<64>(qkv, mask, out, probs, B, g, dc, st) : fwd_d<32>(qkv, mask, out, probs, B, g, dc, st);
}
// dbias (optional): += column sums of dqkv, i.e. the gradient of in_proj_bias
int mha_bwd_mma(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B, int Tn, int H,
                int d, DropCfg dc, cudaStream_t st) {
  const MmaGeom g = make_geom(Tn, H, d);
  MMER_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                 "mha_bwd: qkv/dout/dqkv must be 16-byte aligned");
  if (d == 64 && !g_debug[MMER_DEBUG_ATT_ROWS]) {
    if (g.S <= 8) return bwd_tma_launch<1, 1>(qkv, mask, dout, dqkv, dbias, B, Tn, H, dc, st);
    if (g.S <= 16) return bwd_tma_launch<1, 2>(qkv, mask, dout, dqkv, dbias, B, Tn, H, dc, st);
    if (g.S <= 24) return bwd_tma_launch<2, 3>(qkv, mask, dout, dqkv, dbias, B, Tn, H, dc, st);
    return bwd_tma_launch<2, 4>(qkv, mask, dout, dqkv, dbias, B, Tn, H, dc, st);
  }
  return d == 64 ? bwd_d<64>(qkv, mask, dout, dqkv, dbias, B, g, dc, st)
                 : bwd_d<32>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
}

}  // namespace mmer
