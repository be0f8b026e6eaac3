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
#include "tc05.cuh"

namespace mmer {

int g_debug[16] = {0};

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
  float* rowsum;  // direct (accumulate) path: += sum_k A[m][k] per output row m
  uint8_t* mask_out;          // staged path: 1 bit per output element, set where the stored value is > 0
  const uint8_t* gate_bits;   // staged path: acc *= bit ? gate_scale : 0 (the mask a forward call wrote)
  float* colsum;              // staged path: += column sums of the stored bf16 tile (rows past M hold zeros)
  long long ldmask;           // bytes per row of either bit matrix (N / 8)
  int tma_store;  // bf16 output leaves through swizzled smem staging + TMA store (coalesced)
  int aux_mode;   // staged path only: 0 none, 1 residual add, 2 ReLU gate; the aux tile arrives by TMA
  DropCfg drop;
};

// MMER_GEMM_PROFILE: per-role cycle accounting printed by CTA 0 (tools/ builds only, never the shipped library)
#ifdef MMER_GEMM_PROFILE
#define PROF_DECL(n) long long prof_t[n] = {0}; long long prof_c = clock64(); (void)prof_c
#define PROF_TICK(i) do { long long _n = clock64(); prof_t[i] += _n - prof_c; prof_c = _n; } while (0)
#else
#define PROF_DECL(n)
#define PROF_TICK(i)
#endif
static constexpr int EPI_WARPS = 8;
#ifndef MMER_EPI_BUFS
#define MMER_EPI_BUFS 2
#endif
static constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;

// BN is the N extent of the accumulator tile.  CG = 1: one CTA computes 128 x BN.  CG = 2: a CTA pair computes
// 256 x BN with cta_group::2 MMAs; each CTA loads its own 128 rows of A and HALF of the B tile (BN/2 rows), so the
// L2 -> shared-memory traffic per MMA drops from 48 KB to 32 KB per CTA and k-block at BN = 256.
template <int BN, int CG, bool STAGED>
struct TileCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = STAGED ? 2 * BN : 512;   // direct kernels keep 16 columns at 256 for the row sums of A
  // staged epilogue: per warp EPI_BUFS 32-row x 64-column bf16 tiles (128 B rows; two of them double-buffer the
  // TMA store that drains them); per column half and tile parity one fp32 bias slice of BN/2 columns
  static constexpr int EPI_TILE = 4096;
  static constexpr int EPI_BUFS = MMER_EPI_BUFS;
  static constexpr int BIAS_BYTES = STAGED ? 2 * 2 * (BN / 2) * 4 : 0;
  static constexpr int EPI_BYTES = STAGED ? EPI_WARPS * EPI_BUFS * EPI_TILE : 0;
  static constexpr int ONES_BYTES = STAGED ? 0 : 2048;   // a 16 x 64 bf16 tile of ones: B operand of the row-sum MMAs
  static constexpr int FIXED_BYTES = EPI_BYTES + BIAS_BYTES + ONES_BYTES + 256 /*barriers*/;
  static constexpr int STAGES_FIT = (232448 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert((2 * STAGES + 4 + EPI_WARPS) * 8 + 8 <= 256, "barrier area");
};

// ReLU and dropout of the staged epilogue on 8 consecutive accumulator columns of one row
__device__ __forceinline__ void epi_math8(const GemmParams& p, bool relu, bool drop, float (&v)[8], long long off) {
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (drop) {
    float f[8];
    drop8(p.drop, (uint64_t)off, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= f[i];
  }
}

// Epilogue specialisation of the staged kernels: EPI < 0 decides everything at run time (any combination);
// EPI >= 0 is a bit set fixed at compile time, which removes the predicated code of the unused features from the
// per-element loop -- the epilogue warps are issue-bound, so instruction count is throughput here.
enum : int { EPI_BIAS = 1, EPI_RELU = 2, EPI_DROP = 4, EPI_MASK = 8, EPI_RES = 16, EPI_GATE = 32, EPI_GBITS = 64, EPI_COLSUM = 128 };

template <int BN, bool A_MN, bool B_MN, int CG, bool STAGED, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  using C = TileCfg<BN, CG, STAGED>;
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment: the dynamic shared window is declared (and checked) so
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mmer gemm_tc: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* epi_smem = smem + C::STAGES * C::STAGE_BYTES;
  float* bias_smem = reinterpret_cast<float*>(epi_smem + C::EPI_BYTES);
  uint8_t* ones_smem = epi_smem;   // direct kernels only (EPI_BYTES == 0 there): 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES + C::BIAS_BYTES + C::ONES_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, [2S+4..2S+12) aux, then tmem slot
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + C::STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * C::STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 2 * C::STAGES + 2);
  const uint32_t bar_aux = smem_u32(bars + 2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * C::STAGES + 4 + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;      // CTA of the pair; rank 0 leads (issues the MMAs)
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // persistent work unit (CTA or CTA pair)
  const int num_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, EPI_WARPS * CG);   // the epilogue warps of both CTAs release the leader's MMA issuer
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(bar_aux + 8 * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (STAGED) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (p.aux_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  const bool want_rowsum = !STAGED && p.rowsum != nullptr;
  const bool f_bias = EPI < 0 ? p.bias != nullptr : (EPI & EPI_BIAS) != 0;
  const bool f_relu = EPI < 0 ? p.relu != 0 : (EPI & EPI_RELU) != 0;
  const bool f_drop = EPI < 0 ? p.drop.thr != 0 : (EPI & EPI_DROP) != 0;
  const bool f_mask = EPI < 0 ? p.mask_out != nullptr : (EPI & EPI_MASK) != 0;
  const int f_aux = EPI < 0 ? p.aux_mode : ((EPI & EPI_RES) ? 1 : (EPI & EPI_GATE) ? 2 : 0);
  const bool f_gbits = EPI < 0 ? p.gate_bits != nullptr : (EPI & EPI_GBITS) != 0;
  const bool f_colsum = EPI < 0 ? p.colsum != nullptr : (EPI & EPI_COLSUM) != 0;
  (void)f_colsum;
  (void)f_bias; (void)f_relu; (void)f_drop; (void)f_mask; (void)f_aux; (void)f_gbits;
  if (want_rowsum) {
    // bf16 ones: with them as the B operand an extra N = 16 MMA per k-step accumulates sum_k A[m][k] -- for a
    // weight-gradient GEMM (A = dY^T) that is the bias gradient of the same Linear, at 1/16 of the tile's MMA time
    for (int i = threadIdx.x; i < C::ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32((const void*)tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  const int total_items = p.num_m * p.num_n * p.splits;   // num_m counts (CG*128)-row blocks

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      PROF_DECL(2);
      for (int item = unit; item < total_items; item += num_units) {
        const int n_blk = item % p.num_n;
        const int t = item / p.num_n;
        const int m_blk = t % p.num_m;
        const int split = t / p.num_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = (m_blk * CG + (int)rank) * BM;            // this CTA's rows of A
        const int n0 = n_blk * BN + (int)rank * C::B_ROWS;       // this CTA's share of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          PROF_TICK(1);
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          PROF_TICK(0);
          uint32_t full = bar_full + 8 * stage;
          if (rank == 0) mbar_expect_tx(full, C::STAGE_BYTES * CG);   // the leader's barrier counts both CTAs' bytes
          if (CG == 2) full = mapa_u32(full, 0);
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
            if (CG == 2) tma_load_2d_cg2(dst, map, full, c0, c1); else tma_load_2d(dst, map, full, c0, c1);
          };
          if (!A_MN) {
            load(sa, &tmA, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(sa + c * (BK * 128), &tmA, m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            load(sb, &tmB, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < C::B_ROWS / 64; ++c) load(sb + c * (BK * 128), &tmB, n0 + c * 64, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
#ifdef MMER_GEMM_PROFILE
      if (blockIdx.x == 0) printf("producer: wait_empty %lld issue %lld\n", prof_t[0], prof_t[1]);
#endif
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major SW128: 64-element column
      // blocks BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO).
      const uint32_t mn_lbo = p.mn_swap ? 1024u : (uint32_t)(BK * 128);
      const uint32_t mn_sbo = p.mn_swap ? (uint32_t)(BK * 128) : 1024u;
      const uint32_t a_lbo = A_MN ? mn_lbo : 16u, a_sbo = A_MN ? mn_sbo : 1024u;
      const uint32_t b_lbo = B_MN ? mn_lbo : 16u, b_sbo = B_MN ? mn_sbo : 1024u;
      const uint32_t a_adv = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;  // descriptor units of 16 B
      const uint32_t b_adv = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t idesc_rs = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((uint32_t)(16 >> 3) << 17) |
                                ((uint32_t)((BM * CG) >> 4) << 24);
      const uint64_t ones_desc = make_smem_desc(smem_u32(ones_smem), 16u, 1024u);
      const int nbuf = (want_rowsum && BN == 256) ? 1 : 2;   // the row sums live where the second accumulator would
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t tphase[2] = {0, 0};
      PROF_DECL(3);
      for (int item = unit; item < total_items; item += num_units) {
        const int split = (item / p.num_n) / p.num_m;
        const int rs_nblk = item % p.num_n;   // the tiles of one row block share the row-sum work: k-block kb goes to
                                               // the tile with n_blk == kb % num_n, every tile adds its partial sums
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        PROF_TICK(2);
        mbar_wait(bar_tempty + 8 * buf, tphase[buf] ^ 1);
        PROF_TICK(0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        bool rs_started = false;
        for (int kb = kb0; kb < kb1; ++kb) {
          PROF_TICK(2);
          mbar_wait(bar_full + 8 * stage, phase);
          PROF_TICK(1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo);
          const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16<CG>(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                          (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (want_rowsum && (kb % p.num_n) == rs_nblk) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16<CG>(tmem_base + 256u, adesc + (uint64_t)(k * a_adv), ones_desc + (uint64_t)(k * 2), idesc_rs,
                            (rs_started || k > 0) ? 1u : 0u);
            rs_started = true;
          }
          umma_commit<CG>(bar_empty + 8 * stage);  // frees the smem stage (in both CTAs) when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(bar_tfull + 8 * buf);  // accumulator complete (published to both CTAs' epilogues)
        tphase[buf] ^= 1;
        if (nbuf == 2) buf ^= 1;
      }
#ifdef MMER_GEMM_PROFILE
      if (blockIdx.x == 0) printf("mma: wait_tmem_empty %lld wait_smem_full %lld issue %lld\n", prof_t[0], prof_t[1], prof_t[2]);
#endif
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    // A warp may only touch the TMEM lane quarter (warp % 4); two warps share each quarter and split the
    // BN accumulator columns in halves.
    constexpr int HALF = BN / 2;
    const int ew = warp - 2;
    const int q = warp & 3;
    const int hsel = ew >> 2;
    // staging tiles are 1024-byte aligned, as the 128B swizzle pattern assumes
    uint8_t* stg = epi_smem + ew * (C::EPI_BUFS * C::EPI_TILE);
    const uint32_t stg_u32 = smem_u32(stg);
    const uint32_t auxbar = bar_aux + 8 * ew;
    uint32_t aux_phase = 0;
    uint32_t it = 0;   // staging-tile parity
    bool aux_prefetched = false;
    const float* bias_rd = nullptr;
    float* bias_s = nullptr;
    float4 bias_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    const bool vec_ok = (p.ldd % 8 == 0);
    const uint32_t tempty_leader = CG == 2 ? mapa_u32(bar_tempty, 0) : bar_tempty;
    const int nbuf = (want_rowsum && BN == 256) ? 1 : 2;
    PROF_DECL(6);
    for (int item = unit; item < total_items; item += num_units) {
      const int n_blk = item % p.num_n;
      const int m_blk = (item / p.num_n) % p.num_m;
      const int row0 = (m_blk * CG + (int)rank) * BM + q * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < p.M;
      if (STAGED) {
        if (f_aux && lane < HALF / 64) {
          // warm L2 with this tile's gate / residual sub-tiles while the MMAs of the tile are still running
          const int pc = n_blk * BN + hsel * HALF + lane * 64;
          if (pc < p.N && row0 < p.M)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&tmX), "r"(pc), "r"(row0)
                         : "memory");
        }
        if (f_bias) {
          // the bias slice of this warp's column half (HALF <= 128 values: one float4 per lane) is fetched before
          // the wait for the accumulator and parked in shared memory after it (read back as broadcasts; zero beyond
          // N).  The four warps of a column half write identical values into the same slot; slots alternate with
          // the accumulator buffer.
          bias_s = bias_smem + (buf * 2 + hsel) * HALF;
          const int c = lane * 4;
          if (c < HALF) {
            const int col = n_blk * BN + hsel * HALF + c;
            if (col + 4 <= p.N) {
              bias_reg = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            } else {
              bias_reg = make_float4(0.f, 0.f, 0.f, 0.f);
              if (col < p.N) bias_reg.x = __ldg(p.bias + col);
              if (col + 1 < p.N) bias_reg.y = __ldg(p.bias + col + 1);
              if (col + 2 < p.N) bias_reg.z = __ldg(p.bias + col + 2);
            }
          }
        }
      }
      PROF_TICK(5);
      mbar_wait(bar_tfull + 8 * buf, tphase[buf]);
      PROF_TICK(0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + hsel * HALF);
      if (STAGED && f_bias) {
        // the accumulator of this buffer is ready => every warp has left the tile that last used this bias slot
        if (lane * 4 < HALF) *reinterpret_cast<float4*>(bias_s + lane * 4) = bias_reg;
        bias_rd = bias_s;
        __syncwarp();
      }
      if (STAGED) {
        // ---- bf16 output: registers -> 128B-swizzled smem tile (32 rows x 64 columns) -> TMA store
#pragma unroll 1
        for (int j = 0; j < HALF / 64; ++j) {
          const int col0 = n_blk * BN + hsel * HALF + j * 64;
          if (col0 >= p.N) break;
          const uint32_t tsel = C::EPI_BUFS == 2 ? (it & 1u) : 0u;
          uint8_t* tile = stg + tsel * C::EPI_TILE;
          const uint32_t tile_u32 = stg_u32 + tsel * C::EPI_TILE;
          ++it;
          if (lane == 0 && !aux_prefetched) {
            // the last store that read this staging tile has drained it
            if (C::EPI_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (f_aux) {
              mbar_expect_tx(auxbar, 4096);
              tma_load_2d(tile_u32, &tmX, auxbar, col0, row0);
            }
          }
          aux_prefetched = false;
          __syncwarp();
          PROF_TICK(1);
          // one 64-bit word = this row's gate bits for the 64 columns of the sub-tile (N is a multiple of 64 here)
          uint2 gbits = make_uint2(0u, 0u);
          if (f_gbits && row_ok)
            gbits = __ldg(reinterpret_cast<const uint2*>(p.gate_bits + row * p.ldmask + (col0 >> 3)));
          uint2 mbits = make_uint2(0u, 0u);
          uint32_t r[2][32];
          tmem_ld32_nowait(trow + (uint32_t)(j * 64), r[0]);
          tmem_ld32_nowait(trow + (uint32_t)(j * 64 + 32), r[1]);
          tmem_ld_wait();
          PROF_TICK(2);
          if (f_aux) {
            mbar_wait(auxbar, aux_phase);
            aux_phase ^= 1;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // bias values of the 32 columns first (broadcast reads), so that no shared-memory load has to be ordered
            // behind the staging-tile stores below
            float4 bv[8];
            if (f_bias) {
#pragma unroll
              for (int g = 0; g < 8; ++g) bv[g] = *reinterpret_cast<const float4*>(bias_rd + j * 64 + h * 32 + g * 4);
            }
            uint4 xr[4];
            if (f_aux) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                xr[g] = *reinterpret_cast<const uint4*>(tile + lane * 128 + (((h * 4 + g) ^ (lane & 7)) << 4));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + h * 32 + g * 8;
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[h][g * 8 + i]);
              if (f_bias) {
                v[0] += bv[2 * g].x; v[1] += bv[2 * g].y; v[2] += bv[2 * g].z; v[3] += bv[2 * g].w;
                v[4] += bv[2 * g + 1].x; v[5] += bv[2 * g + 1].y; v[6] += bv[2 * g + 1].z; v[7] += bv[2 * g + 1].w;
              }
              epi_math8(p, f_relu, f_drop, v, row * p.ldd + col);
              if (f_aux) {
                const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr[g]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 xf = __bfloat1622float2(xh[i]);
                  if (f_aux == 1) {
                    v[2 * i] += xf.x;
                    v[2 * i + 1] += xf.y;
                  } else {
                    v[2 * i] *= (xf.x > 0.f) ? p.gate_scale : 0.f;
                    v[2 * i + 1] *= (xf.y > 0.f) ? p.gate_scale : 0.f;
                  }
                }
              }
              if (f_gbits) {
                const uint32_t byte = ((h == 0 ? gbits.x : gbits.y) >> (8 * g)) & 0xFFu;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= ((byte >> i) & 1u) ? p.gate_scale : 0.f;
              }
              if (f_mask) {
                uint32_t byte = 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) byte |= (v[i] > 0.f ? 1u : 0u) << i;
                if (h == 0) mbits.x |= byte << (8 * g); else mbits.y |= byte << (8 * g);
              }
              uint4 pk;
              __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) hp[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(tile + lane * 128 + (((h * 4 + g) ^ (lane & 7)) << 4)) = pk;
            }
          }
          if (f_mask && row_ok)
            *reinterpret_cast<uint2*>(p.mask_out + row * p.ldmask + (col0 >> 3)) = mbits;
          if (f_colsum) {
            // column sums of the tile just staged: lane l owns columns 2l, 2l+1 and walks the 32 rows (one 4-byte word
            // per row; the 32 lanes read the 32 distinct words of a 128-byte row: conflict free).  Rows past M are zero
            // (zero-filled A rows, no bias on this path), so no row predicate is needed.
            __syncwarp();
            float c0 = 0.f, c1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const uint32_t wv = *reinterpret_cast<const uint32_t*>(tile + rr * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)rr & 7u)) << 4) +
                                                                     ((uint32_t)lane & 3u) * 4u);
              c0 += __uint_as_float(wv << 16);
              c1 += __uint_as_float(wv & 0xFFFF0000u);
            }
            const int cc = col0 + 2 * lane;
            if (cc < p.N) atomicAdd(p.colsum + cc, c0);
            if (cc + 1 < p.N) atomicAdd(p.colsum + cc + 1, c1);
          }
          PROF_TICK(3);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmD, tile_u32, col0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (C::EPI_BUFS == 2 && f_aux) {
            // fetch the gate / residual tile of the NEXT sub-tile into the other staging tile now, so that its
            // latency is covered by this sub-tile's store and the next accumulator wait
            int ncol = col0 + 64, nrow = row0;
            bool have = (j + 1 < HALF / 64) && ncol < p.N;
            if (!have) {
              const int nitem = item + num_units;
              if (nitem < total_items) {
                ncol = (nitem % p.num_n) * BN + hsel * HALF;
                nrow = (((nitem / p.num_n) % p.num_m) * CG + (int)rank) * BM + q * 32;
                have = ncol < p.N;
              }
            }
            if (have) {
              if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store before this one has drained
                mbar_expect_tx(auxbar, 4096);
                tma_load_2d(stg_u32 + (it & 1u) * C::EPI_TILE, &tmX, auxbar, ncol, nrow);
              }
              aux_prefetched = true;
            }
          }
          PROF_TICK(4);
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
      bool rs_any = false;
      if (want_rowsum) {
        const int split = (item / p.num_n) / p.num_m;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        // first k-block >= kb0 that is congruent to n_blk modulo num_n
        const int first = kb0 + ((n_blk - kb0 % p.num_n) + p.num_n) % p.num_n;
        rs_any = first < kb1;
      }
      if (rs_any && hsel == 0) {
        uint32_t rs;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(rs) : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + 256u));
        tmem_ld_wait();
        if (row_ok) atomicAdd(p.rowsum + row, __uint_as_float(rs));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader + 8 * buf); else mbar_arrive(bar_tempty + 8 * buf);
      }
      tphase[buf] ^= 1;
      if (nbuf == 2) buf ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef MMER_GEMM_PROFILE
    if (blockIdx.x == 0 && lane == 0)
      printf("epi warp %d: wait_tmem_full %lld wait_store_drain %lld tmem_ld %lld math %lld fence+store %lld other %lld\n", ew,
             prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4], prof_t[5]);
#endif
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // no CTA of a pair leaves while its peer may still signal it
  if (warp == 1) tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
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
int make_tma_map_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1) {
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

template <int BN, bool A_MN, bool B_MN, int CG, bool STAGED, int EPI = -1>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx,
                  const GemmParams& p, int grid, cudaStream_t st) {
  using C = TileCfg<BN, CG, STAGED>;
  static unsigned long long attr_done = 0ull;   // one bit per device (cudaFuncSetAttribute is per device)
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CG, STAGED, EPI>;
  if (needs_func_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc)");
  }
  cudaError_t e = launch_dep(kern, dim3((unsigned)grid), dim3(GEMM_THREADS), C::SMEM_BYTES, st, CG, ta, tb, td, tx, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_tc)");
  MMER_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

// true when gemm_tc sends this problem's output through the staged (shared-memory + TMA store) epilogue
bool gemm_tc_stages_output(const mmer_gemm_args& a) {
  return a.out_dtype == MMER_BF16 && !a.accumulate && a.ldd % 8 == 0 && !(a.gate && a.residual) && !g_debug[MMER_DEBUG_DIRECT_STORE];
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
  MMER_CHECK_ARG((a.relu_mask_out == nullptr && a.gate_bits == nullptr) ||
                     (a.out_dtype == MMER_BF16 && !a.accumulate && a.N % 64 == 0 && a.ldd == a.N && a.gate == nullptr),
                 "gemm_tc: bit masks need a dense bf16 output with N %% 64 == 0 and no gate tensor");
  MMER_CHECK_ARG(a.a_rowsum == nullptr || (a.accumulate && a.a_major == MMER_MAJOR_MN),
                 "gemm_tc: a_rowsum needs accumulate mode and an MN-major A (weight-gradient GEMM)");
  MMER_CHECK_ARG(!(a.a_major == MMER_MAJOR_MN && a.b_major == MMER_MAJOR_K), "gemm_tc: (MN,K) operand majors unused");
  MMER_CHECK_ARG(a.d_colsum == nullptr || (a.bias == nullptr && a.residual == nullptr && !a.accumulate),
                 "gemm_tc: d_colsum is for data-gradient GEMMs (no bias, no residual, no accumulation)");

  const int nsm = sm_count();
  // CTA pairs (cta_group::2, 256 x 256 tiles) whenever the problem offers enough of them to fill the machine;
  // otherwise single CTAs with 128 x 256 or 128 x 128 tiles.
  const int pairs = nsm / 2;
  int cg = 1;
  int bn = 256;
  {
    const long long pair_tiles = (long long)ceil_div(a.M, 2 * BM) * ceil_div(a.N, 256);
    const long long kb = ceil_div(a.K, BK);
    const bool enough = a.accumulate ? (pair_tiles * (kb / 8 > 0 ? kb / 8 : 1) >= pairs / 2) : (pair_tiles >= pairs);
    if (a.M >= 2 * BM && a.N > 128 && enough && g_debug[MMER_DEBUG_NO_PAIR] == 0) cg = 2;
  }
  const int bm_eff = BM * cg;
  const int num_m = ceil_div(a.M, bm_eff);
  if (cg == 1) {
    if ((long long)num_m * ceil_div(a.N, 256) < nsm && a.N > 128 && !a.accumulate) bn = 128;
    if (a.N <= 128) bn = 128;
    if (g_debug[MMER_DEBUG_FORCE_BN] == 128 || g_debug[MMER_DEBUG_FORCE_BN] == 256) bn = g_debug[MMER_DEBUG_FORCE_BN];
  }
  const int units = cg == 2 ? pairs : nsm;
  const int num_n = ceil_div(a.N, bn);
  const int kb_total = ceil_div(a.K, BK);
  int splits = 1;
  if (a.accumulate) {
    const int tiles = num_m * num_n;
    int want = units / tiles;
    if (want < 1) want = 1;
    const int max_by_k = kb_total / 8 > 0 ? kb_total / 8 : 1;  // at least 8 k-blocks per split
    if (want > max_by_k) want = max_by_k;
    splits = want;
    if (g_debug[MMER_DEBUG_FORCE_SPLITS] > 0 && g_debug[MMER_DEBUG_FORCE_SPLITS] <= max_by_k) splits = g_debug[MMER_DEBUG_FORCE_SPLITS];
  }
  int kb_per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per);

  CUtensorMap ta, tb;
  if (a.a_major == MMER_MAJOR_K) {
    MMER_TRY(make_tma_map_bf16(&ta, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, 64, BM));
  } else {
    MMER_TRY(make_tma_map_bf16(&ta, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, 64, BK));
  }
  if (a.b_major == MMER_MAJOR_K) {
    MMER_TRY(make_tma_map_bf16(&tb, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, 64, (uint32_t)(bn / cg)));
  } else {
    MMER_TRY(make_tma_map_bf16(&tb, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, 64, BK));
  }

  CUtensorMap td = ta, tx = ta;  // placeholders when the staged store / aux tile are not used
  const bool tma_store = a.out_dtype == MMER_BF16 && !a.accumulate && a.ldd % 8 == 0 && !(a.gate && a.residual) &&
                         !g_debug[MMER_DEBUG_DIRECT_STORE];
  const void* aux = a.gate ? a.gate : a.residual;
  if (tma_store) {
    MMER_CHECK_ARG(aux == nullptr || (reinterpret_cast<uintptr_t>(aux) & 15) == 0,
                   "gemm_tc: gate/residual must be 16-byte aligned");
    MMER_TRY(make_tma_map_bf16(&td, a.D, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldd, 64, 32));
    if (aux) MMER_TRY(make_tma_map_bf16(&tx, aux, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldd, 64, 32));
  }

  GemmParams p;
  p.tma_store = tma_store ? 1 : 0;
  p.aux_mode = !tma_store ? 0 : (a.gate ? 2 : (a.residual ? 1 : 0));
  p.M = (int)a.M; p.N = (int)a.N; p.K = (int)a.K;
  p.num_m = num_m; p.num_n = num_n; p.splits = splits; p.kb_total = kb_total; p.kb_per_split = kb_per;
  p.rowsum = a.a_rowsum;
  p.mask_out = a.relu_mask_out; p.gate_bits = a.gate_bits; p.ldmask = a.N / 8;
  p.colsum = tma_store ? a.d_colsum : nullptr;   // other output paths: a separate pass (gemm_dispatch)
  p.D = a.D; p.ldd = a.ldd; p.bias = a.bias; p.residual = a.residual; p.gate = a.gate; p.gate_scale = a.gate_scale;
  p.out_f32 = a.out_dtype == MMER_F32; p.accumulate = a.accumulate; p.relu = a.relu;
  p.mn_swap = g_debug[MMER_DEBUG_MN_SWAP];
  p.drop = make_drop(a.drop_p, a.seed, a.drop_site);
  long long items = (long long)num_m * num_n * splits;
  int grid = (int)(items < units ? items : units) * cg;

  const bool amn = a.a_major == MMER_MAJOR_MN, bmn = a.b_major == MMER_MAJOR_MN;
#define MMER_GEMM_LAUNCH(BN_, CG_)                                                                             \
  do {                                                                                                         \
    if (tma_store) {                                                                                           \
      if (!amn && !bmn) return launch<BN_, false, false, CG_, true>(ta, tb, td, tx, p, grid, st);              \
      if (!amn && bmn) return launch<BN_, false, true, CG_, true>(ta, tb, td, tx, p, grid, st);                \
      return launch<BN_, true, true, CG_, true>(ta, tb, td, tx, p, grid, st);                                  \
    }                                                                                                          \
    if (!amn && !bmn) return launch<BN_, false, false, CG_, false>(ta, tb, td, tx, p, grid, st);               \
    if (!amn && bmn) return launch<BN_, false, true, CG_, false>(ta, tb, td, tx, p, grid, st);                 \
    return launch<BN_, true, true, CG_, false>(ta, tb, td, tx, p, grid, st);                                   \
  } while (0)
  if (cg == 2 && tma_store && g_debug[MMER_DEBUG_GENERIC_EPI] == 0) {
    // the step's hot epilogues get their own instantiation (see EPI_* above); anything else runs the generic one
    const int mode = (a.bias ? EPI_BIAS : 0) | (a.relu ? EPI_RELU : 0) | (p.drop.thr ? EPI_DROP : 0) |
                     (a.relu_mask_out ? EPI_MASK : 0) | (p.aux_mode == 1 ? EPI_RES : 0) | (p.aux_mode == 2 ? EPI_GATE : 0) |
                     (a.gate_bits ? EPI_GBITS : 0) | (p.colsum ? EPI_COLSUM : 0);
    if (!amn && !bmn) {
      if (mode == EPI_BIAS) return launch<256, false, false, 2, true, EPI_BIAS>(ta, tb, td, tx, p, grid, st);
      if (mode == (EPI_BIAS | EPI_RELU))   // linear1 in eval mode (inference forward)
        return launch<256, false, false, 2, true, EPI_BIAS | EPI_RELU>(ta, tb, td, tx, p, grid, st);
      if (mode == (EPI_BIAS | EPI_RELU | EPI_DROP | EPI_MASK))
        return launch<256, false, false, 2, true, EPI_BIAS | EPI_RELU | EPI_DROP | EPI_MASK>(ta, tb, td, tx, p, grid, st);
      // z = residual + dropout(x W^T + b): the pre-LayerNorm sum of a post-norm sub-layer, written instead of the
      // sub-layer output (engine.cu, ln_mode 2)
      if (mode == (EPI_BIAS | EPI_DROP | EPI_RES))
        return launch<256, false, false, 2, true, EPI_BIAS | EPI_DROP | EPI_RES>(ta, tb, td, tx, p, grid, st);
      if (mode == (EPI_BIAS | EPI_RES))
        return launch<256, false, false, 2, true, EPI_BIAS | EPI_RES>(ta, tb, td, tx, p, grid, st);
    } else if (!amn && bmn) {
      if (mode == 0) return launch<256, false, true, 2, true, 0>(ta, tb, td, tx, p, grid, st);
      if (mode == EPI_RES) return launch<256, false, true, 2, true, EPI_RES>(ta, tb, td, tx, p, grid, st);
      if (mode == EPI_GBITS) return launch<256, false, true, 2, true, EPI_GBITS>(ta, tb, td, tx, p, grid, st);
      if (mode == (EPI_GBITS | EPI_COLSUM))   // linear2 dgrad + linear1 bias gradient
        return launch<256, false, true, 2, true, EPI_GBITS | EPI_COLSUM>(ta, tb, td, tx, p, grid, st);
    }
  }
  if (cg == 2) MMER_GEMM_LAUNCH(256, 2);
  if (bn == 256) MMER_GEMM_LAUNCH(256, 1);
  MMER_GEMM_LAUNCH(128, 1);
#undef MMER_GEMM_LAUNCH
}

}  // namespace mmer
