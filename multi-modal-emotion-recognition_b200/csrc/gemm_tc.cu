// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] . B[N,K]^T), bf16 in, fp32 accumulate.
//
// One persistent CTA per SM, 10 warps:
//   warp 0   : TMA producer (one elected lane) -- cp.async.bulk.tensor into a 4-stage smem ring,
//              128B-swizzled tiles, completion on "full" mbarriers
//   warp 1   : MMA issuer (one elected lane) -- tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16,
//              accumulators in TMEM (2 x BN columns, double buffered); tcgen05.commit releases smem
//              stages and publishes finished accumulators.  Also owns tcgen05.alloc / dealloc.
//   warps 2-9: epilogue -- tcgen05.ld (32 lanes x 32 columns per warp; two warps per TMEM lane quarter, each
//              taking half of the BN columns), bias / ReLU / dropout in registers; the gate or residual tile
//              arrives by TMA (L2-prefetched while the tile's MMAs run); bf16 results leave through a
//              128B-swizzled smem tile and a TMA store.  fp32 outputs and split-K weight gradients
//              (fp32 vector atomics) use direct stores.
// Either operand may be K-major (nn.Linear forward, dgrad's dY) or MN-major (dgrad's W, wgrad's dY and X),
// which is what lets dgrad and wgrad run without any transposed copies in HBM.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace mmer {

static constexpr int BM = 128;
static constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;

int g_debug[16] = {0};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor, SWIZZLE_128B (layout type 2), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct GemmParams {
  int M, N, K;
  int num_m, num_n, splits, kb_total, kb_per_split;
  void* D;
  long long ldd;
  const float* bias;
  const void* residual;
  const void* gate;
  float gate_scale;
  int out_f32, accumulate, relu;
  int mn_swap;    // debug: swap LBO/SBO of MN-major descriptors
  int tma_store;  // bf16 output leaves through swizzled smem staging + TMA store (coalesced)
  int aux_mode;   // staged path only: 0 none, 1 residual add, 2 ReLU gate; the aux tile arrives by TMA
  DropCfg drop;
};

static constexpr int EPI_WARPS = 8;
static constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;

template <int BN>
struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_BYTES = EPI_WARPS * 4096;  // per warp: one 32-row x 64-column bf16 tile (128 B rows)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert((2 * STAGES + 4 + EPI_WARPS) * 8 + 8 <= 256, "barrier area");
};

// element-wise part of the epilogue on 8 consecutive accumulator columns of one row
__device__ __forceinline__ void epi_math8(const GemmParams& p, float (&v)[8], int col, long long off, bool full8) {
  if (p.bias) {
    if (full8) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.N) v[i] += __ldg(p.bias + col + i);
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.drop.thr) {
    float f[8];
    drop8(p.drop, (uint64_t)off, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= f[i];
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  using C = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, [2S+4..2S+12) aux, then tmem slot
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + C::STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * C::STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 2 * C::STAGES + 2);
  const uint32_t bar_aux = smem_u32(bars + 2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * C::STAGES + 4 + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, EPI_WARPS);
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(bar_aux + 8 * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (p.aux_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = p.num_m * p.num_n * p.splits;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int n_blk = item % p.num_n;
        const int t = item / p.num_n;
        const int m_blk = t % p.num_m;
        const int split = t / p.num_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, C::STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, full, kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmA, full, m_blk * BM + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, full, kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmB, full, n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major SW128: 64-element column
      // blocks BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO).
      const uint32_t mn_lbo = p.mn_swap ? 1024u : (uint32_t)(BK * 128);
      const uint32_t mn_sbo = p.mn_swap ? (uint32_t)(BK * 128) : 1024u;
      const uint32_t a_lbo = A_MN ? mn_lbo : 16u, a_sbo = A_MN ? mn_sbo : 1024u;
      const uint32_t b_lbo = B_MN ? mn_lbo : 16u, b_sbo = B_MN ? mn_sbo : 1024u;
      const uint32_t a_adv = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;  // descriptor units of 16 B
      const uint32_t b_adv = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase[2] = {0, 0};
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = (item / p.num_n) / p.num_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(bar_tempty + 8 * buf, tphase[buf] ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo);
          const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);  // frees the smem stage when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * buf);  // accumulator complete
        tphase[buf] ^= 1;
        buf ^= 1;
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    // A warp may only touch the TMEM lane quarter (warp % 4); two warps share each quarter and split the
    // BN accumulator columns in halves.
    constexpr int HALF = BN / 2;
    const int ew = warp - 2;
    const int q = warp & 3;
    const int hsel = ew >> 2;
    uint8_t* stg = epi_smem + ew * 4096;
    const uint32_t stg_u32 = smem_u32(stg);
    const uint32_t auxbar = bar_aux + 8 * ew;
    uint32_t aux_phase = 0;
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    const bool vec_ok = (p.ldd % 8 == 0);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int n_blk = item % p.num_n;
      const int m_blk = (item / p.num_n) % p.num_m;
      const int row0 = m_blk * BM + q * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < p.M;
      if (p.aux_mode && lane < HALF / 64) {
        // warm L2 with this tile's gate / residual sub-tiles while the MMAs of the tile are still running
        const int pc = n_blk * BN + hsel * HALF + lane * 64;
        if (pc < p.N && row0 < p.M)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tmX), "r"(pc), "r"(row0)
                       : "memory");
      }
      mbar_wait(bar_tfull + 8 * buf, tphase[buf]);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + hsel * HALF);
      if (p.tma_store) {
        // ---- bf16 output: registers -> 128B-swizzled smem tile (32 rows x 64 columns) -> TMA store
#pragma unroll 1
        for (int j = 0; j < HALF / 64; ++j) {
          const int col0 = n_blk * BN + hsel * HALF + j * 64;
          if (col0 >= p.N) break;
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has drained the tile
            if (p.aux_mode) {
              mbar_expect_tx(auxbar, 4096);
              tma_load_2d(stg_u32, &tmX, auxbar, col0, row0);
            }
          }
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(trow + (uint32_t)(j * 64 + h * 32), r);
            if (p.aux_mode && h == 0) {
              mbar_wait(auxbar, aux_phase);
              aux_phase ^= 1;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + h * 32 + g * 8;
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
              epi_math8(p, v, col, row * p.ldd + col, col + 8 <= p.N);
              uint4* slot = reinterpret_cast<uint4*>(stg + lane * 128 + (((h * 4 + g) ^ (lane & 7)) << 4));
              if (p.aux_mode) {
                const uint4 xr = *slot;
                const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 xf = __bfloat1622float2(xh[i]);
                  if (p.aux_mode == 1) {
                    v[2 * i] += xf.x;
                    v[2 * i + 1] += xf.y;
                  } else {
                    v[2 * i] *= (xf.x > 0.f) ? p.gate_scale : 0.f;
                    v[2 * i + 1] *= (xf.y > 0.f) ? p.gate_scale : 0.f;
                  }
                }
              }
              uint4 pk;
              __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) hp[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              *slot = pk;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmD, stg_u32, col0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
        // ---- direct path: fp32 outputs, split-K accumulation (fp32 vector atomics), unaligned leading dims
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)c, r);
          const int col0 = n_blk * BN + hsel * HALF + c;
          if (!row_ok || col0 >= p.N) continue;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col >= p.N) break;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            const bool full8 = vec_ok && (col + 8 <= p.N);
            const int nvalid = min(8, p.N - col);
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nvalid) v[i] += __ldg(p.bias + col + i);
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            const long long off = row * p.ldd + col;
            if (p.drop.thr && full8) {
              float f[8];
              drop8(p.drop, (uint64_t)off, f);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= f[i];
            } else if (p.drop.thr) {
              for (int i = 0; i < nvalid; ++i) v[i] *= drop1(p.drop, (uint64_t)(off + i));
            }
            if (p.out_f32) {
              float* D = reinterpret_cast<float*>(p.D) + off;
              if (p.gate) {
                const float* G = reinterpret_cast<const float*>(p.gate) + off;
                for (int i = 0; i < nvalid; ++i) v[i] *= (G[i] > 0.f) ? p.gate_scale : 0.f;
              }
              if (p.residual) {
                const float* R = reinterpret_cast<const float*>(p.residual) + off;
                for (int i = 0; i < nvalid; ++i) v[i] += R[i];
              }
              if (p.accumulate) {
                if (full8) {
                  atomicAdd(reinterpret_cast<float4*>(D), make_float4(v[0], v[1], v[2], v[3]));
                  atomicAdd(reinterpret_cast<float4*>(D + 4), make_float4(v[4], v[5], v[6], v[7]));
                } else {
                  for (int i = 0; i < nvalid; ++i) atomicAdd(D + i, v[i]);
                }
              } else if (full8) {
                store8(D, v);
              } else {
                for (int i = 0; i < nvalid; ++i) D[i] = v[i];
              }
            } else {
              bf16* D = reinterpret_cast<bf16*>(p.D) + off;
              if (full8) {
                if (p.gate) {
                  float gv[8];
                  load8(reinterpret_cast<const bf16*>(p.gate) + off, gv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] *= (gv[i] > 0.f) ? p.gate_scale : 0.f;
                }
                if (p.residual) {
                  float rv[8];
                  load8(reinterpret_cast<const bf16*>(p.residual) + off, rv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += rv[i];
                }
                store8(D, v);
              } else {
                for (int i = 0; i < nvalid; ++i) {
                  float x = v[i];
                  if (p.gate) x *= (to_f(reinterpret_cast<const bf16*>(p.gate)[off + i]) > 0.f) ? p.gate_scale : 0.f;
                  if (p.residual) x += to_f(reinterpret_cast<const bf16*>(p.residual)[off + i]);
                  D[i] = __float2bfloat16_rn(x);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      tphase[buf] ^= 1;
      buf ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ k.d0; h = h * 1000003u ^ k.d1; h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ k.b0; h = h * 1000003u ^ k.b1;
    return h;
  }
};

// 2-D bf16 tensor map: inner dimension d0 (contiguous), outer d1 with row stride ld elements.
static int make_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, d0, d1, ld, b0, b1};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return MMER_ERR_CUDA; }
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%llu,%llu ld=%llu box=%u,%u", (int)r, ptr,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    return MMER_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
                  const GemmParams& p, int grid, cudaStream_t st) {
  using C = TileCfg<BN>;
  static bool attr_done = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc)");
    attr_done = true;
  }
  kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, st>>>(ta, tb, td, tx, p);
  MMER_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

int gemm_tc(const mmer_gemm_args& a, cudaStream_t st) {
  MMER_CHECK_ARG(a.in_dtype == MMER_BF16, "gemm_tc: bf16 inputs only");
  MMER_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm_tc: empty problem M=%lld N=%lld K=%lld", (long long)a.M,
                 (long long)a.N, (long long)a.K);
  MMER_CHECK_ARG(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm_tc: lda/ldb must be multiples of 8 elements (16 B)");
  MMER_CHECK_ARG((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a.D) & 15) == 0,
                 "gemm_tc: pointers must be 16-byte aligned");
  MMER_CHECK_ARG(!a.accumulate || a.out_dtype == MMER_F32, "gemm_tc: accumulate needs fp32 output");
  MMER_CHECK_ARG(!(a.a_major == MMER_MAJOR_MN && a.b_major == MMER_MAJOR_K), "gemm_tc: (MN,K) operand majors unused");

  const int nsm = sm_count();
  const int num_m = ceil_div(a.M, BM);
  int bn = 256;
  if ((long long)num_m * ceil_div(a.N, 256) < nsm && a.N > 128 && !a.accumulate) bn = 128;
  if (a.N <= 128) bn = 128;
  if (g_debug[MMER_DEBUG_FORCE_BN] == 128 || g_debug[MMER_DEBUG_FORCE_BN] == 256) bn = g_debug[MMER_DEBUG_FORCE_BN];
  const int num_n = ceil_div(a.N, bn);
  const int kb_total = ceil_div(a.K, BK);
  int splits = 1;
  if (a.accumulate) {
    const int tiles = num_m * num_n;
    int want = nsm / tiles;
    if (want < 1) want = 1;
    const int max_by_k = kb_total / 8 > 0 ? kb_total / 8 : 1;  // at least 8 k-blocks per split
    if (want > max_by_k) want = max_by_k;
    splits = want;
  }
  int kb_per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per);

  CUtensorMap ta, tb;
  if (a.a_major == MMER_MAJOR_K) {
    MMER_TRY(make_map(&ta, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, 64, BM));
  } else {
    MMER_TRY(make_map(&ta, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, 64, BK));
  }
  if (a.b_major == MMER_MAJOR_K) {
    MMER_TRY(make_map(&tb, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, 64, (uint32_t)bn));
  } else {
    MMER_TRY(make_map(&tb, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, 64, BK));
  }

  CUtensorMap td = ta, tx = ta;  // placeholders when the staged store / aux tile are not used
  const bool tma_store = a.out_dtype == MMER_BF16 && !a.accumulate && a.ldd % 8 == 0 && !(a.gate && a.residual) &&
                         !g_debug[MMER_DEBUG_DIRECT_STORE];
  const void* aux = a.gate ? a.gate : a.residual;
  if (tma_store) {
    MMER_CHECK_ARG(aux == nullptr || (reinterpret_cast<uintptr_t>(aux) & 15) == 0,
                   "gemm_tc: gate/residual must be 16-byte aligned");
    MMER_TRY(make_map(&td, a.D, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldd, 64, 32));
    if (aux) MMER_TRY(make_map(&tx, aux, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldd, 64, 32));
  }

  GemmParams p;
  p.tma_store = tma_store ? 1 : 0;
  p.aux_mode = !tma_store ? 0 : (a.gate ? 2 : (a.residual ? 1 : 0));
  p.M = (int)a.M; p.N = (int)a.N; p.K = (int)a.K;
  p.num_m = num_m; p.num_n = num_n; p.splits = splits; p.kb_total = kb_total; p.kb_per_split = kb_per;
  p.D = a.D; p.ldd = a.ldd; p.bias = a.bias; p.residual = a.residual; p.gate = a.gate; p.gate_scale = a.gate_scale;
  p.out_f32 = a.out_dtype == MMER_F32; p.accumulate = a.accumulate; p.relu = a.relu;
  p.mn_swap = g_debug[MMER_DEBUG_MN_SWAP];
  p.drop = make_drop(a.drop_p, a.seed, a.drop_site);
  long long items = (long long)num_m * num_n * splits;
  int grid = (int)(items < nsm ? items : nsm);

  const bool amn = a.a_major == MMER_MAJOR_MN, bmn = a.b_major == MMER_MAJOR_MN;
  if (bn == 256) {
    if (!amn && !bmn) return launch<256, false, false>(ta, tb, td, tx, p, grid, st);
    if (!amn && bmn) return launch<256, false, true>(ta, tb, td, tx, p, grid, st);
    return launch<256, true, true>(ta, tb, td, tx, p, grid, st);
  } else {
    if (!amn && !bmn) return launch<128, false, false>(ta, tb, td, tx, p, grid, st);
    if (!amn && bmn) return launch<128, false, true>(ta, tb, td, tx, p, grid, st);
    return launch<128, true, true>(ta, tb, td, tx, p, grid, st);
  }
}

}  // namespace mmer
