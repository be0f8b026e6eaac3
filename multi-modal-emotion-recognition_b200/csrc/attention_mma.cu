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
//             dK = dS^T Q; the transposed operands (Pd^T, dS^T) go through a 2.5 KB per-warp scratch tile and
//             ldmatrix.trans
// Results overwrite operand slots that are dead by then (O -> Q slot; dV -> V slot, dK -> K slot, dQ -> dO slot),
// so whole token rows leave through bulk shared->global copies.  HBM traffic is the algorithmic minimum: every
// input byte is read once, every output byte written once, all as >= 1 KB contiguous bursts.
// Rows/keys beyond S are handled by clamping fragment addresses to row S-1 (finite data) and zeroing their
// probabilities, so no shared memory beyond the S real rows is needed.
#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

static constexpr int MMA_WARPS = 8;
static constexpr int SCR_STRIDE = 80;                 // bytes per scratch row (32 bf16 + 16 B pad: conflict-free ldmatrix)
static constexpr int SCR_BYTES = 32 * SCR_STRIDE;     // per warp

struct MmaGeom {
  int S, F, Tn, H;
  uint32_t in_row, in_stride;   // bytes of one packed qkv row, padded smem stride
  uint32_t do_row, do_stride;   // bytes of one dO / out row, padded smem stride
};

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
template <int D, int MT, int NT>
__device__ __forceinline__ void scores_softmax(uint32_t qbase, uint32_t kbase, uint32_t stride, int S, int lane,
                                               uint32_t kvalid, float (&p)[MT][NT][4]) {
  constexpr int KS = D / 16;
  uint32_t kf[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int k2 = 0; k2 < KS / 2; ++k2) {
      const int row = min(nt * 8 + (lane & 7), S - 1);
      const int col = k2 * 32 + (lane >> 3) * 8;
      ldsm_x4(kbase + row * stride + col * 2, kf[nt][2 * k2][0], kf[nt][2 * k2][1], kf[nt][2 * k2 + 1][0],
              kf[nt][2 * k2 + 1][1]);
    }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) p[mt][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a0, a1, a2, a3;
      const int row = min(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S - 1);
      const int col = ks * 16 + (lane >> 4) * 8;
      ldsm_x4(qbase + row * stride + col * 2, a0, a1, a2, a3);
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
template <int D, int KSTEPS>
__device__ __forceinline__ void mma_rows_x(float (&acc)[D / 8][4], const uint32_t (&a)[KSTEPS][4], uint32_t xbase,
                                           uint32_t stride, int S, int lane) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
    for (int n2 = 0; n2 < D / 16; ++n2) {
      uint32_t b0, b1, b2, b3;
      const int row = min(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S - 1);
      const int col = n2 * 16 + (lane >> 4) * 8;
      ldsm_x4_t(xbase + row * stride + col * 2, b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * n2], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b0, b1);
      mma_bf16_16816(acc[2 * n2 + 1], a[ks][0], a[ks][1], a[ks][2], a[ks][3], b2, b3);
    }
}

// store a 16 x D accumulator tile as bf16 rows (row0 + g, row0 + g + 8) of a smem matrix, rows >= S skipped
template <int D>
__device__ __forceinline__ void store_tile(uint8_t* base, uint32_t stride, int row0, int S, int lane,
                                           const float (&acc)[D / 8][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + 8 * r;
    if (row < S) {
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
        *reinterpret_cast<uint32_t*>(base + (size_t)row * stride + (nd * 8 + t * 2) * 2) =
            pack_bf16x2(acc[nd][2 * r], acc[nd][2 * r + 1]);
    }
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
    float p[MT][NT][4];
    scores_softmax<D, MT, NT>(qbase, kbase, gm.in_stride, S, lane, kvalid, p);
    const long long bh = (long long)b * H + h;
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
                drop2(dc, att_drop_index(bh * S + i, j, NT * 8), f0, f1);
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
      mma_rows_x<D, (NT + 1) / 2>(o, pa[mt], vbase, gm.in_stride, S, lane);
      store_tile<D>(smem + h * D * 2, gm.in_stride, mt * 16, S, lane, o);   // O_h overwrites the dead Q_h slot
    }
  }
  fence_async_smem();
  __syncthreads();
  if (warp == 0) {
    for (int r = lane; r < S; r += 32) bulk_s2g(out + ((long long)b * S + r) * F, in_a + r * gm.in_stride, gm.do_row);
    bulk_commit();
    bulk_wait_read0();
  }
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
  uint8_t* scr = do_s + do_bytes + warp * SCR_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(do_s + do_bytes + MMA_WARPS * SCR_BYTES);
  float* colacc = reinterpret_cast<float*>(bar + 2);   // [3F] running column sums of dqkv (in_proj bias gradient)
  const uint32_t bar_a = smem_u32(bar), in_a = smem_u32(smem), do_a = smem_u32(do_s), scr_a = smem_u32(scr);
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
    const long long bh = (long long)b * H + h;
    float p[MT][NT][4];
    scores_softmax<D, MT, NT>(qbase, kbase, gm.in_stride, S, lane, kvalid, p);
    // dP = dO V^T
    float dp[MT][NT][4];
    {
      uint32_t vf[NT][KS][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int k2 = 0; k2 < KS / 2; ++k2) {
          const int row = min(nt * 8 + (lane & 7), S - 1);
          const int col = k2 * 32 + (lane >> 3) * 8;
          ldsm_x4(vbase + row * gm.in_stride + col * 2, vf[nt][2 * k2][0], vf[nt][2 * k2][1], vf[nt][2 * k2 + 1][0],
                  vf[nt][2 * k2 + 1][1]);
        }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) dp[mt][nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a0, a1, a2, a3;
          const int row = min(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, S - 1);
          const int col = ks * 16 + (lane >> 4) * 8;
          ldsm_x4(dobase + row * gm.do_stride + col * 2, a0, a1, a2, a3);
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
          if (dc.thr && row_ok) drop2(dc, att_drop_index(bh * S + i, nt * 8 + t * 2, NT * 8), f[nt][0], f[nt][1]);
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
    // ---- dV = Pd^T dO  (Pd^T through the scratch tile)
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          *reinterpret_cast<uint32_t*>(scr + (mt * 16 + g + 8 * r) * SCR_STRIDE + (nt * 8 + t * 2) * 2) =
              pack_bf16x2(p[mt][nt][r * 2], p[mt][nt][r * 2 + 1]);
    __syncwarp();
#pragma unroll
    for (int mk = 0; mk < MTK; ++mk) {
      uint32_t a[MT][4];
#pragma unroll
      for (int kq = 0; kq < MT; ++kq) {
        const int row = kq * 16 + (lane & 7) + (lane >> 4) * 8;         // query (reduction index)
        const int col = mk * 16 + ((lane >> 3) & 1) * 8;                // key (output row)
        ldsm_x4_t(scr_a + row * SCR_STRIDE + col * 2, a[kq][0], a[kq][1], a[kq][2], a[kq][3]);
      }
      float acc[D / 8][4];
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nd][i] = 0.f;
      mma_rows_x<D, MT>(acc, a, dobase, gm.do_stride, S, lane);
      store_tile<D>(smem + 2 * F * 2 + h * D * 2, gm.in_stride, mk * 16, S, lane, acc);   // dV_h -> dead V_h slot
    }
    // ---- dS^T through the same scratch tile (for dK); dS fragments stay in registers (for dQ)
    uint32_t dsa[MT][(NT + 1) / 2][4];
    pack_rows<MT, NT>(dp, dsa);
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          *reinterpret_cast<uint32_t*>(scr + (mt * 16 + g + 8 * r) * SCR_STRIDE + (nt * 8 + t * 2) * 2) =
              pack_bf16x2(dp[mt][nt][r * 2], dp[mt][nt][r * 2 + 1]);
    __syncwarp();
    // ---- dQ = dS K -> dead dO_h slot
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float acc[D / 8][4];
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nd][i] = 0.f;
      mma_rows_x<D, (NT + 1) / 2>(acc, dsa[mt], kbase, gm.in_stride, S, lane);
      store_tile<D>(do_s + h * D * 2, gm.do_stride, mt * 16, S, lane, acc);
    }
    // ---- dK = dS^T Q -> dead K_h slot
    float acck[MTK][D / 8][4];
#pragma unroll
    for (int mk = 0; mk < MTK; ++mk) {
      uint32_t a[MT][4];
#pragma unroll
      for (int kq = 0; kq < MT; ++kq) {
        const int row = kq * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = mk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(scr_a + row * SCR_STRIDE + col * 2, a[kq][0], a[kq][1], a[kq][2], a[kq][3]);
      }
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
        for (int i = 0; i < 4; ++i) acck[mk][nd][i] = 0.f;
      mma_rows_x<D, MT>(acck[mk], a, qbase, gm.in_stride, S, lane);
    }
    __syncwarp();   // every lane has finished reading K_h (dQ) before it is overwritten
#pragma unroll
    for (int mk = 0; mk < MTK; ++mk)
      store_tile<D>(smem + F * 2 + h * D * 2, gm.in_stride, mk * 16, S, lane, acck[mk]);
    __syncwarp();
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
         (size_t)MMA_WARPS * SCR_BYTES + 16 + (size_t)3 * g.F * sizeof(float);
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
  return d == 64 ? fwd_d<64>(qkv, mask, out, probs, B, g, dc, st) : fwd_d<32>(qkv, mask, out, probs, B, g, dc, st);
}
// dbias (optional): += column sums of dqkv, i.e. the gradient of in_proj_bias
int mha_bwd_mma(const void* qkv, const uint8_t* mask, const void* dout, void* dqkv, float* dbias, int B, int Tn, int H,
                int d, DropCfg dc, cudaStream_t st) {
  const MmaGeom g = make_geom(Tn, H, d);
  MMER_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                 "mha_bwd: qkv/dout/dqkv must be 16-byte aligned");
  return d == 64 ? bwd_d<64>(qkv, mask, dout, dqkv, dbias, B, g, dc, st)
                 : bwd_d<32>(qkv, mask, dout, dqkv, dbias, B, g, dc, st);
}

}  // namespace mmer
