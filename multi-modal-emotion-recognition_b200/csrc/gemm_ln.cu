// Linear -> (+bias) -> dropout -> (+residual) -> LayerNorm in ONE tcgen05 kernel, for the post-norm encoder sub-layers
// (nn.TransformerEncoderLayer, norm_first=False, as configured at train2.py:111-118 / train.py:54-57):
//
//     z = x + dropout(a W^T + b)          y = LayerNorm(z) * gamma + beta          N = d_model = 512
//
// replaces a GEMM launch + add_ln_fwd_pipe_kernel (out_proj -> norm1 and linear2 -> norm2 of every layer).
//
// Structure = the CTA-pair GEMM of gemm_tc.cu (TMA producer warp, single-thread tcgen05.mma cta_group::2 issuer, fp32
// accumulators double-buffered in the 512 TMEM columns) with a different tile order and two groups of epilogue warps.
// A LayerNorm row needs all 512 output columns, i.e. BOTH 256-column accumulator tiles of a row block, so a CTA pair
// walks row blocks (256 rows) and computes their two n-tiles back to back into the two TMEM buffers:
//   warps 2-9  (pass 1, the GEMM epilogue): tcgen05.ld (32 columns at a time) -> + bias -> dropout -> + residual tile
//           (prefetched by TMA) -> z; per-row shifted sums (pivot = the row's first value) accumulate in registers;
//           bf16(z) leaves through a 128B-swizzled staging tile as a TMA store into the Z tensor (which the backward
//           kernel reads instead of re-reading x and a).  The TMEM buffer is released right after its last tcgen05.ld.
//           The two warps that share a TMEM lane quarter hold 256 columns each of the same 32 rows: they swap
//           (pivot, S1, S2) through shared memory behind a 64-thread named barrier and merge them (Chan), which gives
//           mean / rstd without cancellation; (mean, rstd) go to shared memory and to the stats tensor for backward.
//   warps 10-17 (pass 2, normalisers): once a row block's z tiles are complete in global memory (they are L2-resident,
//           written microseconds ago) these warps stream them back with coalesced 16-byte loads -- a warp per row, a
//           lane per 8 columns, gamma / beta in registers -- and write y = LN(bf16 z) with coalesced 16-byte stores.
//           y is computed from exactly the bf16 z that backward will see.  No TMEM, no shared-memory tiles: their
//           instructions fill the issue slots the latency-bound pass-1 warps leave idle, one row block behind them.
// No S x 512 fp32 row ever lives in registers or shared memory, and the pipeline keeps 4 operand stages.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tc05.cuh"

namespace mmer {

namespace {

constexpr int LN_N = 512;            // d_model: two 256-column accumulator tiles
constexpr int LN_BN = 256;
constexpr int LN_EPI_WARPS = 8;      // pass-1 warps (TMEM readers): 2 per lane quarter, 128 of an n-tile's 256 columns each
constexpr int LN_NRM_WARPS = 8;      // pass-2 warps (normalisers)
constexpr int LN_THREADS = 64 + 32 * (LN_EPI_WARPS + LN_NRM_WARPS);
constexpr int LN_A_BYTES = BM * BK * 2;                 // 16 KB: this CTA's 128 rows of A
constexpr int LN_B_BYTES = (LN_BN / 2) * BK * 2;        // 16 KB: this CTA's half of the 256-row B tile
constexpr int LN_STAGE_BYTES = LN_A_BYTES + LN_B_BYTES;
constexpr int LN_STAGES = 4;
constexpr int LN_TILE = 4096;                           // 32 rows x 64 bf16 columns, 128B-swizzled
constexpr int LN_EPI_BYTES = LN_EPI_WARPS * 2 * LN_TILE; // two staging tiles per pass-1 warp
constexpr int LN_VEC_BYTES = 3 * LN_N * 4;              // bias, gamma, beta
constexpr int LN_XCH_BYTES = LN_EPI_WARPS * 32 * 16;    // (pivot, S1, S2, pad) per lane
constexpr int LN_RST_BYTES = 2 * BM * 8;                // (mean, rstd) of the CTA's 128 rows, two row blocks in flight
constexpr int LN_BAR_BYTES = 512;
constexpr int LN_SMEM_BYTES =
    LN_STAGES * LN_STAGE_BYTES + LN_EPI_BYTES + LN_VEC_BYTES + LN_XCH_BYTES + LN_RST_BYTES + LN_BAR_BYTES;
static_assert(LN_SMEM_BYTES <= 232448, "shared memory budget");
static_assert((2 * LN_STAGES + 4 + 2 * LN_EPI_WARPS + 4) * 8 + 8 <= LN_BAR_BYTES, "barrier area");
constexpr float LN_EPS = 1e-5f;

struct GemmLnParams {
  int M, K;
  int num_rb;          // row blocks of 256 rows
  int kb_total;        // k-blocks of 64
  const float* bias;   // [512] or NULL
  const float* gamma;  // [512]
  const float* beta;   // [512]
  float* stats;        // [M][2] (mean, rstd)
  const bf16* Z;       // [M][512]: written by pass 1 through tmZ, read back by pass 2
  bf16* Y;             // [M][512]
  int has_residual;
  int dbg;             // MMER_DEBUG_LN_VARIANT bits (A/B timing only): 1 no residual tile, 2 no z store, 4 no pass 2, 8 no epilogue math
  DropCfg drop;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(LN_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ, const GemmLnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mmer gemm_ln: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* epi_smem = smem + LN_STAGES * LN_STAGE_BYTES;
  float* vec_smem = reinterpret_cast<float*>(epi_smem + LN_EPI_BYTES);          // bias | gamma | beta
  float4* xch_smem = reinterpret_cast<float4*>(epi_smem + LN_EPI_BYTES + LN_VEC_BYTES);
  float2* rst_smem = reinterpret_cast<float2*>(epi_smem + LN_EPI_BYTES + LN_VEC_BYTES + LN_XCH_BYTES);   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + LN_EPI_BYTES + LN_VEC_BYTES + LN_XCH_BYTES + LN_RST_BYTES);
  // bars: [0..S) full, [S..2S) empty, 2 tmem_full, 2 tmem_empty, 16 aux (one per residual staging tile), 2 z_ready, 2 z_done, tmem slot
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + LN_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * LN_STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 2 * LN_STAGES + 2);
  const uint32_t bar_aux = smem_u32(bars + 2 * LN_STAGES + 4);
  const uint32_t bar_zready = smem_u32(bars + 2 * LN_STAGES + 4 + 2 * LN_EPI_WARPS);
  const uint32_t bar_zdone = smem_u32(bars + 2 * LN_STAGES + 4 + 2 * LN_EPI_WARPS + 2);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * LN_STAGES + 4 + 2 * LN_EPI_WARPS + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)(blockIdx.x >> 1);
  const int num_units = (int)(gridDim.x >> 1);

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < LN_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, LN_EPI_WARPS * 2);   // the pass-1 warps of both CTAs release the leader's issuer
      mbar_init(bar_zready + 8 * b, LN_EPI_WARPS);
      mbar_init(bar_zdone + 8 * b, LN_NRM_WARPS);
    }
    for (int w = 0; w < 2 * LN_EPI_WARPS; ++w) mbar_init(bar_aux + 8 * w, 1);   // one per staging tile
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
  }
  if (warp == 1) tmem_alloc<2>(smem_u32((const void*)tmem_slot), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int rb = unit; rb < p.num_rb; rb += num_units) {
        const int m0 = (rb * 2 + (int)rank) * BM;
        for (int n = 0; n < 2; ++n) {
          const int n0 = n * LN_BN + (int)rank * (LN_BN / 2);
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            uint32_t full = bar_full + 8 * stage;
            if (rank == 0) mbar_expect_tx(full, LN_STAGE_BYTES * 2);   // the leader's barrier counts both CTAs' bytes
            full = mapa_u32(full, 0);
            const uint32_t sa = smem_u32(smem + stage * LN_STAGE_BYTES);
            tma_load_2d_cg2(sa, &tmA, full, kb * BK, m0);
            tma_load_2d_cg2(sa + LN_A_BYTES, &tmB, full, kb * BK, n0);
            if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA, one lane)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LN_BN >> 3) << 17) | ((uint32_t)((BM * 2) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tphase[2] = {0, 0};
      for (int rb = unit; rb < p.num_rb; rb += num_units) {
        for (int n = 0; n < 2; ++n) {      // n-tile n lives in TMEM buffer n
          mbar_wait(bar_tempty + 8 * n, tphase[n] ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(n * LN_BN);
          for (int kb = 0; kb < p.kb_total; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * LN_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa, 16u, 1024u);
            const uint64_t bdesc = make_smem_desc(sa + LN_A_BYTES, 16u, 1024u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16<2>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit<2>(bar_empty + 8 * stage);
            if (++stage == LN_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit<2>(bar_tfull + 8 * n);
          tphase[n] ^= 1;
        }
      }
    }
  } else {
    // bias / gamma / beta: 6 KB of shared memory filled by the 16 epilogue warps
    for (int i = threadIdx.x - 64; i < LN_N; i += 32 * (LN_EPI_WARPS + LN_NRM_WARPS)) {
      vec_smem[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
      vec_smem[LN_N + i] = __ldg(p.gamma + i);
      vec_smem[2 * LN_N + i] = __ldg(p.beta + i);
    }
    named_bar_sync(1, 32 * (LN_EPI_WARPS + LN_NRM_WARPS));
    if (warp < 2 + LN_EPI_WARPS) {
      // ===================================================== pass 1 (warps 2..9): z tiles + row statistics
      const int ew = warp - 2;
      const int q = warp & 3;          // TMEM lane quarter this warp may read
      const int hsel = ew >> 2;        // which 128 of the n-tile's 256 columns
      const float* bias_s = vec_smem;
      uint8_t* stg = epi_smem + ew * (2 * LN_TILE);
      const uint32_t stg_u32 = smem_u32(stg);
      const uint32_t auxbar = bar_aux + 16 * ew;    // two barriers: one per staging tile
      uint32_t aux_phase[2] = {0, 0};
      uint32_t it = 0;                  // staging-tile parity
      uint32_t tphase[2] = {0, 0};
      const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
      const bool has_res = p.has_residual != 0 && !(p.dbg & 1);
      int iter = 0;                     // row blocks done by this CTA
      bool pending = false;             // z_ready of the previous row block not signalled yet
      bool aux_ready = false;           // the residual tile of the CURRENT sub-tile has already been requested
      for (int rb = unit; rb < p.num_rb; rb += num_units, ++iter) {
        const int par = iter & 1;
        const int row0 = (rb * 2 + (int)rank) * BM + q * 32;
        const long long row = (long long)row0 + lane;
        const bool block_ok = row0 < p.M;      // rows past M: TMA clips the stores, the statistics are not written
        float pivot = 0.f, s1 = 0.f, s2 = 0.f;
        if (has_res && lane < 4) {
          // warm L2 with this warp's four residual tiles of the CTA's NEXT row block
          const int nrow0 = ((rb + num_units) * 2 + (int)rank) * BM + q * 32;
          if (nrow0 < p.M)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tmX),
                         "r"((lane >> 1) * LN_BN + hsel * 128 + (lane & 1) * 64), "r"(nrow0)
                         : "memory");
        }
#pragma unroll 1
        for (int n = 0; n < 2; ++n) {
          mbar_wait(bar_tfull + 8 * n, tphase[n]);
          tphase[n] ^= 1;
          tc_fence_after();
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(n * LN_BN + hsel * 128);
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            const int col0 = n * LN_BN + hsel * 128 + j * 64;
            const uint32_t tsel = it & 1u;
            uint8_t* tile = stg + tsel * LN_TILE;
            const uint32_t tile_u32 = stg_u32 + tsel * LN_TILE;
            ++it;
            if (lane == 0) {
              // Residual tiles are requested a whole sub-tile ahead: at the top of sub-tile s the tile of s + 1 goes out
              // into the OTHER staging tile (whose last reader, the z store of s - 1, has had a sub-tile's time to drain),
              // so that a TMA load's ~1 us of latency is never waited for.
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              if (has_res && block_ok && !aux_ready) {            // the very first sub-tile of this warp
                mbar_expect_tx(auxbar + 8 * tsel, LN_TILE);
                tma_load_2d(tile_u32, &tmX, auxbar + 8 * tsel, col0, row0);
              }
              const bool last = (n == 1 && j == 1);
              const int nrb = last ? rb + num_units : rb;
              const int nrow0 = (nrb * 2 + (int)rank) * BM + q * 32;
              const int ncol = last ? hsel * 128 : (j == 0 ? col0 + 64 : LN_BN + hsel * 128);
              if (has_res && nrb < p.num_rb && nrow0 < p.M) {
                mbar_expect_tx(auxbar + 8 * (tsel ^ 1u), LN_TILE);
                tma_load_2d(stg_u32 + (tsel ^ 1u) * LN_TILE, &tmX, auxbar + 8 * (tsel ^ 1u), ncol, nrow0);
              }
            }
            {
              const bool last = (n == 1 && j == 1);
              const int nrb = last ? rb + num_units : rb;
              aux_ready = has_res && nrb < p.num_rb && (nrb * 2 + (int)rank) * BM + q * 32 < p.M;
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t r[32];
              tmem_ld32(trow + (uint32_t)(j * 64 + h * 32), r);
              if (j == 1 && h == 1) {
                // the whole accumulator buffer has been read: hand it back to the MMA issuer now
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * n);
              }
              if (h == 0 && has_res && block_ok) {
                mbar_wait(auxbar + 8 * tsel, aux_phase[tsel]);
                aux_phase[tsel] ^= 1;
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (p.dbg & 8) break;
                const int col = col0 + h * 32 + g * 8;
                uint8_t* addr = tile + lane * 128 + (((h * 4 + g) ^ (lane & 7)) << 4);
                const float4 b0 = *reinterpret_cast<const float4*>(bias_s + col);
                const float4 b1 = *reinterpret_cast<const float4*>(bias_s + col + 4);
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
                if (has_res) {
                  const uint4 xr = *reinterpret_cast<const uint4*>(addr);
                  const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 xf = __bfloat1622float2(xh[i]);
                    x[2 * i] = xf.x;
                    x[2 * i + 1] = xf.y;
                  }
                }
                if (DROP) {
                  float f[8];
                  drop8(p.drop, (uint64_t)(row * LN_N + col), f);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = fmaf(f[i], v[i], x[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += x[i];
                }
                if (n == 0 && j == 0 && h == 0 && g == 0) pivot = v[0];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float d = v[i] - pivot;
                  s1 += d;
                  s2 = fmaf(d, d, s2);
                }
                uint4 pk;
                __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) hp[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                *reinterpret_cast<uint4*>(addr) = pk;
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && block_ok && !(p.dbg & 2)) {
              tma_store_2d(&tmZ, tile_u32, col0, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (pending && n == 0 && j == 0) {
              // every z store of the PREVIOUS row block has now been followed by a newer group: once at most one group
              // is pending they are complete in global memory -> let the normaliser warps at that row block
              if (lane == 0) {
                if (block_ok) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // no newer group was committed
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_arrive(bar_zready + 8 * (par ^ 1));
              }
              pending = false;
            }
          }
        }
        // ---------------------------------------------------------------- row statistics (two warps per 32 rows)
        xch_smem[ew * 32 + lane] = make_float4(pivot, s1, s2, 0.f);
        // the (mean, rstd) slot of this parity was last read by the normalisers two row blocks ago
        if (iter >= 2) mbar_wait(bar_zdone + 8 * par, (uint32_t)((iter >> 1) - 1) & 1u);
        named_bar_sync(2 + q, 64);
        const float4 o = xch_smem[(ew ^ 4) * 32 + lane];
        const float inv_h = 1.f / 256.f;
        const float mean_a = pivot + s1 * inv_h, m2_a = s2 - s1 * s1 * inv_h;
        const float mean_b = o.x + o.y * inv_h, m2_b = o.z - o.y * o.y * inv_h;
        const float delta = mean_b - mean_a;
        const float mean = 0.5f * (mean_a + mean_b);
        const float var = (m2_a + m2_b + delta * delta * 128.f) * (1.f / (float)LN_N);
        const float rstd = rsqrtf(fmaxf(var, 0.f) + LN_EPS);
        if (hsel == 0) {
          rst_smem[par * BM + q * 32 + lane] = make_float2(mean, rstd);
          if (row < p.M) *reinterpret_cast<float2*>(p.stats + row * 2) = make_float2(mean, rstd);
        }
        named_bar_sync(2 + q, 64);                       // xch slot reusable; (mean, rstd) written before either warp signals
        pending = true;
      }
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (pending) {
          asm volatile("fence.proxy.async;" ::: "memory");
          mbar_arrive(bar_zready + 8 * ((iter - 1) & 1));
        }
      }
    } else {
      // ===================================================== pass 2 (warps 10..17): y = LN(bf16 z), global -> global
      const int pw = warp - 2 - LN_EPI_WARPS;
      constexpr int ROWS = BM / LN_NRM_WARPS;    // 16 rows of the CTA's 128 per warp
      float gm[2][8], bt[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gm[c][i] = vec_smem[LN_N + c * 256 + lane * 8 + i];
          bt[c][i] = vec_smem[2 * LN_N + c * 256 + lane * 8 + i];
        }
      int iter = 0;
      for (int rb = unit; rb < p.num_rb; rb += num_units, ++iter) {
        const int par = iter & 1;
        mbar_wait(bar_zready + 8 * par, (uint32_t)(iter >> 1) & 1u);
        const long long base = (long long)(rb * 2 + (int)rank) * BM + pw * ROWS;
#pragma unroll 1
        for (int r0 = 0; r0 < ROWS; r0 += 4) {
          if (p.dbg & 4) break;
          uint4 zr[4][2];
          float2 st[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const long long row = base + r0 + k;
            st[k] = rst_smem[par * BM + pw * ROWS + r0 + k];
            if (row < p.M) {
              const uint4* src = reinterpret_cast<const uint4*>(p.Z + row * LN_N);
              zr[k][0] = __ldcg(src + lane);
              zr[k][1] = __ldcg(src + 32 + lane);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const long long row = base + r0 + k;
            if (row >= p.M) continue;
            const float rstd = st[k].y, nmr = -st[k].x * st[k].y;
            uint4* dst = reinterpret_cast<uint4*>(p.Y + row * LN_N);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const __nv_bfloat162* zh = reinterpret_cast<const __nv_bfloat162*>(&zr[k][c]);
              uint4 pk;
              __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 zf = __bfloat1622float2(zh[i]);
                hp[i] = __floats2bfloat162_rn(fmaf(fmaf(zf.x, rstd, nmr), gm[c][2 * i], bt[c][2 * i]),
                                              fmaf(fmaf(zf.y, rstd, nmr), gm[c][2 * i + 1], bt[c][2 * i + 1]));
              }
              dst[c * 32 + lane] = pk;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_zdone + 8 * par);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace

int gemm_ln_fwd(const void* A, const void* W, const float* bias, const void* residual, const float* gamma, const float* beta,
                void* z_out, void* y_out, float* stats, long long M, long long K, DropCfg drop, cudaStream_t st) {
  MMER_CHECK_ARG(A && W && gamma && beta && z_out && y_out && stats, "gemm_ln_fwd: null pointer");
  MMER_CHECK_ARG(M > 0 && K > 0 && K % 8 == 0, "gemm_ln_fwd: M > 0 and K a positive multiple of 8 (M=%lld K=%lld)", M, K);
  MMER_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(z_out) |
                   reinterpret_cast<uintptr_t>(y_out) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0,
                 "gemm_ln_fwd: pointers must be 16-byte aligned");
  CUtensorMap ta, tb, tx, tz;
  MMER_TRY(make_tma_map_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)K, 64, BM));
  MMER_TRY(make_tma_map_bf16(&tb, W, (uint64_t)K, (uint64_t)LN_N, (uint64_t)K, 64, LN_BN / 2));
  MMER_TRY(make_tma_map_bf16(&tz, z_out, (uint64_t)LN_N, (uint64_t)M, (uint64_t)LN_N, 64, 32));
  tx = tz;
  if (residual) MMER_TRY(make_tma_map_bf16(&tx, residual, (uint64_t)LN_N, (uint64_t)M, (uint64_t)LN_N, 64, 32));
  GemmLnParams p;
  p.M = (int)M; p.K = (int)K;
  p.num_rb = ceil_div(M, 2 * BM);
  p.kb_total = ceil_div(K, BK);
  p.bias = bias; p.gamma = gamma; p.beta = beta; p.stats = stats;
  p.Z = reinterpret_cast<const bf16*>(z_out); p.Y = reinterpret_cast<bf16*>(y_out);
  p.has_residual = residual != nullptr;
  p.dbg = g_debug[MMER_DEBUG_LN_VARIANT];
  p.drop = drop;
  const int pairs = sm_count() / 2;
  const int grid = (p.num_rb < pairs ? p.num_rb : pairs) * 2;
  static unsigned long long attr_done[2] = {0ull, 0ull};
  cudaError_t e;
  if (drop.thr) {
    auto kern = gemm_ln_kernel<true>;
    if (needs_func_attr(&attr_done[0])) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM_BYTES);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_ln)");
    }
    e = launch_dep(kern, dim3((unsigned)grid), dim3(LN_THREADS), LN_SMEM_BYTES, st, 2, ta, tb, tx, tz, p);
  } else {
    auto kern = gemm_ln_kernel<false>;
    if (needs_func_attr(&attr_done[1])) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM_BYTES);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_ln)");
    }
    e = launch_dep(kern, dim3((unsigned)grid), dim3(LN_THREADS), LN_SMEM_BYTES, st, 2, ta, tb, tx, tz, p);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_ln)");
  MMER_LAUNCH_CHECK("gemm_ln_kernel");
  return 0;
}

}  // namespace mmer

using namespace mmer;

extern "C" int mmer_gemm_ln_fwd(const void* a, const void* w, const float* bias, const void* residual, const float* gamma,
                                const float* beta, void* z_out, void* y_out, float* stats, int64_t M, int64_t K,
                                float drop_p, uint32_t site, uint64_t seed, void* stream) {
  return gemm_ln_fwd(a, w, bias, residual, gamma, beta, z_out, y_out, stats, (long long)M, (long long)K,
                     make_drop(drop_p, seed, site), (cudaStream_t)stream);
}
