// tcgen05 / TMEM / TMA GEMM for the Conformer hot path (sm_100a only).
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )      bf16 operands, fp32 accumulation in TMEM
//
// Operand layouts (what the "major" flags mean):
//   a_mn = 0 : A is row-major [M][K] (K contiguous, "K-major")        -- activations in the forward pass
//   a_mn = 1 : A is stored     [K][M] (M contiguous, "MN-major")       -- dY in a weight-gradient GEMM
//   b_mn = 0 : B is row-major [N][K] (torch Linear weight (out,in))    -- forward
//   b_mn = 1 : B is stored     [K][N] (N contiguous)                   -- W in dgrad, X in wgrad
// so forward = (0,0), dgrad dX = dY*W = (0,1), wgrad dW = dY^T*X = (1,1): no transposes are ever materialised.
//
// Replaces, on the reference path: cuBLASLt GEMMs behind torch.nn.Linear / Conv1d(k=1) in
// torchaudio/models/conformer.py:44-51,66-73,102-109 and the in/out projections of nn.MultiheadAttention
// (torch/nn/functional.py:5798,6690), with the bias / SiLU / 0.5*residual / dropout elementwise kernels fused as epilogue.
//
// Structure: PERSISTENT kernel, one CTA per SM, 576 threads, static round-robin over 128 x BN output tiles
// (n fastest, so CTAs running side by side share the same A rows in L2):
//   warp 0     TMA producer (one elected lane) -> 128B-swizzled smem ring (mbarrier full/empty), runs ahead across tiles
//   warp 1     TMEM allocator + MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16,
//              accumulators DOUBLE-BUFFERED in TMEM (2 x BN columns): tile i+1's main loop overlaps tile i's epilogue
//   warps 2-17 epilogue (16 warps; warp w owns TMEM lanes 32*(w%4).. and one quarter of the BN columns), two families:
//              ROW DOMAIN (bf16 outputs of the plain / SiLU epilogues): tcgen05.ld (thread = row) -> bias / SiLU / dropout on
//              the thread's 32 consecutive columns -> bf16 -> 64B-swizzled 2 KB box in smem (a second box for the saved
//              pre-activation) -> TMA store (clips at M / N).  Half the smem traffic of a transpose, no LDS / STG stream.
//              COALESCED DOMAIN (fp32 residual outputs, SiLU-backward with its aux operand and fused bias-gradient column
//              sums, split-K accumulation): tcgen05.ld -> release the TMEM buffer -> XOR-swizzled per-warp fp32 transpose ->
//              lanes along columns: 128-byte coalesced global loads & stores (or red.global.add.v4.f32); the residual /
//              aux operands are fetched with L1::no_allocate loads before the accumulator is awaited.
// The problems here are skinny (K = 144..2304) so a tile's main loop is only 3-16 k-blocks and the kernel is bound by
// SHARED-MEMORY BANDWIDTH (operand fills + MMA operand reads + epilogue staging; measured with in-kernel cycle stamps,
// DESIGN.md section 4), not by the tensor pipe: hence the overlap through double-buffered TMEM, compile-time epilogue
// variants, MUFU.TANH-based SiLU, and a tile schedule whose last, mostly empty wave runs as half-width tiles.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int BS_MAXKB = 4;                   // weight-stationary path: k-blocks of the B tile kept resident (K <= 256)
#ifndef B2C_DEFAULT_CLUSTER
#define B2C_DEFAULT_CLUSTER 1                   // multicast cluster size picked when block_n = 0 (env B2C_GEMM_CLUSTER overrides)
#endif

#ifdef B2C_GEMM_TIMING
// debug builds only (make EXTRA=-DB2C_GEMM_TIMING): cycle stamps of CTA 0, read back with b2c_debug_gemm_timing
__device__ long long g_gemm_dbg[3 * 16 * 32];
#define B2C_GT(role, it_, slot) do { if (blockIdx.x == 0 && (it_) < 16 && (slot) < 32) g_gemm_dbg[((role) * 16 + (it_)) * 32 + (slot)] = clock64(); } while (0)
#else
#define B2C_GT(role, it_, slot) do {} while (0)
#endif

struct GemmParams {
  DropCfg drop;
  int64_t M, N, K;
  int kblocks_per_split, total_kb;
  int m_tiles, n_tiles, splits, total_tiles;
  int tiles_full;            // tiles [tiles_full, total_tiles) are HALF-width tail tiles (two per full tile); == total_tiles when unused
  int bs_group, bs_rounds, bs_rem, bs_half;   // weight-stationary schedule (see tile_at)
  b2c_epilogue ep;
};

// BSTAT = weight-stationary variant: the CTA keeps ONE n-block of B (all K <= 256 of it) resident in shared memory for
// its whole life and the ring streams A only.  With 128 KB of B parked, a 256-wide tile leaves room for 8 epilogue warps.
// AUX3 (weight-stationary SiLU-backward dgrad on 128-wide tiles): three 2 KB boxes per epilogue warp instead of two.
template <int BN, bool CTA2 = false, bool BSTAT = false, bool AUX3 = false>
struct GemmCfg {
  static constexpr int BNH = CTA2 ? BN / 2 : BN;                         // B rows held by ONE CTA
  static constexpr int EPW = (BSTAT && BN == 256) ? 8 : 16;              // epilogue warps: 4 TMEM lane quarters x EPW/4 column slices
  static constexpr int THREADS = 64 + EPW * 32;
  static constexpr int STAGES = BSTAT ? (BN == 256 ? 3 : (AUX3 ? 4 : 5)) : ((BNH == 256) ? 3 : (BNH == 128 ? 4 : 6));
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BNH * BK * 2;
  static constexpr uint32_t STAGE_BYTES = BSTAT ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr uint32_t BRES_OFF = STAGES * STAGE_BYTES;             // resident B: BS_MAXKB k-blocks of B_BYTES
  static constexpr uint32_t BRES_BYTES = BSTAT ? BS_MAXKB * B_BYTES : 0;
  static constexpr int SPAN = BN / (EPW / 4);                            // columns owned by one epilogue warp
  static constexpr int CW = SPAN < 32 ? SPAN : 32;                       // columns per epilogue chunk (one tcgen05.ld)
  static constexpr int CHUNKS = SPAN / CW;                               // chunks per epilogue warp per tile
  static constexpr int AUX_BUFS = AUX3 ? 3 : 2;                          // 2 KB boxes per warp in the SiLU-backward epilogue
  static constexpr uint32_t STG_WARP_BYTES = AUX3 ? 3 * 2048 : 32 * CW * 4;   // fp32 transpose buffer / bf16 boxes per warp
  static constexpr uint32_t STG_OFF = BRES_OFF + BRES_BYTES;
  static constexpr uint32_t BAR_OFF = STG_OFF + EPW * STG_WARP_BYTES;
  static constexpr uint32_t AUXBAR_OFF = BAR_OFF + (AUX3 ? 128 : 256);   // AUX_BUFS mbarriers per epilogue warp (aux boxes of the SiLU' epilogue)
  static_assert((2 * STAGES + 6) * 8 <= (AUX3 ? 128 : 256), "pipeline barriers overlap the aux barriers");
  static constexpr uint32_t BIAS_OFF = BAR_OFF + 512;                    // bias copy read by the row-domain epilogues (LDS broadcast)
  static constexpr uint32_t BIAS_MAXN = BSTAT ? BN : 2048;               // stationary CTAs only ever see their own n-block
  static constexpr uint32_t CS_OFF = BIAS_OFF + BIAS_MAXN * 4;           // per-CTA column-sum accumulators (training epilogues)
  static constexpr uint32_t CS_MAXN = BSTAT ? BN : 2048;                 // (stationary: indexed from the CTA's first column)
  static constexpr uint32_t SMEM = CS_OFF + 1024 /*align slack*/;
  static constexpr uint32_t SMEM_EXTRA = SMEM + CS_MAXN * 4;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ float sigmoid_fast(float z) { return sigmoid_f(z); }

// Epilogue specialisations (compile-time, so the per-element code stays a few dozen instructions and fits the
// instruction cache -- the fully generic runtime-switched version was I-cache bound: ncu stall_no_inst dominated).
enum : int {
  EPI_GENERIC = 0,   // everything decided at run time (any combination the ABI allows)
  EPI_BF16 = 1,      // out = bf16(acc [+ bias])                                  QKV / pointwise-1 / plain dgrad
  EPI_SILU = 2,      // [out2 = bf16(acc+bias)]; out = bf16(drop(SiLU(acc+bias)))  FFN first linear
  EPI_RESID = 3,     // out = resid + alpha * drop(acc + bias)      (fp32)          FFN second linear / out-proj / pointwise-2
  EPI_SILU_BWD = 4,  // out = bf16(drop(acc * SiLU'(aux)))                          dgrad through the FFN activation
  EPI_ACCUM = 5,     // out += alpha * acc   (fp32 red.global.add.v4)               weight gradients (split-K)
};

// Tile schedule.  Tiles [0, tiles_full) are full 128 x BN tiles, n fastest.  When the last wave of a launch would be
// less than half full (187 tiles on 148 SMs: the second wave runs at 26 %), its tiles are issued as TWO half-width
// tiles each (hh = 0 / 1), so that wave costs ~0.67 of a full one instead of 1.0 -- single CTAs without split-K only.
struct TileCoord { int m_blk, n_blk, split, hh; };
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int cdiv, uint32_t rank) {
  TileCoord t;
  if (tile < p.tiles_full) {
    t.n_blk = tile % p.n_tiles;
    t.m_blk = ((tile / p.n_tiles) % p.m_tiles) * cdiv + static_cast<int>(rank);
    t.split = tile / (p.n_tiles * p.m_tiles);
    t.hh = -1;
  } else {
    const int t2 = tile - p.tiles_full, b = p.tiles_full + (t2 >> 1);
    t.n_blk = b % p.n_tiles;
    t.m_blk = b / p.n_tiles;
    t.split = 0;
    t.hh = t2 & 1;
  }
  return t;
}

// it-th tile of this CTA.  Default schedule: static round-robin over decode_tile's order.  Weight-stationary schedule:
// CTA c is bound to n-block c % n_tiles for its whole life (that B tile never leaves shared memory) and is member
// j = c / n_tiles of that n-block's group of bs_group CTAs; the group walks the m-blocks round-robin for bs_rounds full
// rounds, and the bs_rem m-blocks left over go out as full tiles or, when they fill at most half the group, as two
// half-width tiles each (bs_half) -- CTAs side by side (same j) still share their A rows in L2.
template <bool BSTAT, int CDIV>
__device__ __forceinline__ bool tile_at(const GemmParams& p, int it, uint32_t rank, TileCoord& t) {
  if constexpr (BSTAT) {
    const int j = static_cast<int>(blockIdx.x) / p.n_tiles;
    t.n_blk = static_cast<int>(blockIdx.x) - j * p.n_tiles;
    t.split = 0;
    t.hh = -1;
    if (it < p.bs_rounds) { t.m_blk = j + it * p.bs_group; return true; }
    if (it > p.bs_rounds) return false;
    if (p.bs_half) {
      if (j >= 2 * p.bs_rem) return false;
      t.m_blk = p.bs_rounds * p.bs_group + (j >> 1);
      t.hh = j & 1;
      return true;
    }
    if (j >= p.bs_rem) return false;
    t.m_blk = p.bs_rounds * p.bs_group + j;
    return true;
  } else {
    const int tile = static_cast<int>(blockIdx.x) / CDIV + it * (static_cast<int>(gridDim.x) / CDIV);
    if (tile >= p.total_tiles) return false;
    t = decode_tile(p, tile, CDIV, rank);
    return true;
  }
}

// CTA2 = true: launched as clusters of two CTAs; tile = 256 x BN per pair (tcgen05.mma.cta_group::2, M = 256).  CTA r of
// the pair owns rows [128 r, 128 r + 128) of the tile (A slab, accumulator, epilogue) and loads B rows
// [r BN/2, (r+1) BN/2); the leader (r = 0) issues every MMA and its commits fan out to both CTAs' barriers.
// EXTRA = false compiles the training-only extras (saved pre-activation, dropout, fused column sums) out of the
// specialised epilogues; the generic epilogue always keeps them (run-time switched).
// CL > 1 (only with CTA2 = false): launched as clusters of CL CTAs that work on CL vertically adjacent 128-row tiles of
// the SAME n-block.  Every CTA fetches 1/CL of the B tile and TMA-multicasts it into all CL shared memories, so the
// B bytes cross the L2 -> SM fabric once per cluster instead of once per CTA (these GEMMs are bound by that fabric,
// not by HBM or the tensor pipe: K is only 256..1024, so a 128 x BN tile re-reads BN x K weights for ~1-2 k cycles of
// MMA).  A stage is recycled only after the MMAs of all CL CTAs that read it have retired (multicast commit).
template <int BN, bool A_MN, bool B_MN, int EPI, bool CTA2, bool EXTRA, int CL = 1, bool BSTAT = false>
__global__ void __launch_bounds__((GemmCfg<BN, CTA2, BSTAT>::THREADS), 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                       const __grid_constant__ CUtensorMap tmB,
                                                                       const __grid_constant__ CUtensorMap tmO,
                                                                       const __grid_constant__ CUtensorMap tmO2,
                                                                       const __grid_constant__ CUtensorMap tmBh,
                                                                       const GemmParams p) {
  pdl_trigger();                 // PDL: the next kernel may be scheduled; pdl_wait() follows the prologue below
  constexpr bool AUX3 = BSTAT && BN == 128 && EPI == EPI_SILU_BWD;
  using Cfg = GemmCfg<BN, CTA2, BSTAT, AUX3>;
  constexpr int BNH = Cfg::BNH;
  constexpr int EPI_WARPS = Cfg::EPW;
  constexpr int GEMM_THREADS = Cfg::THREADS;
  static_assert(!(CTA2 && CL > 1), "multicast clusters and CTA pairs are separate paths");
  static_assert(!BSTAT || (!CTA2 && CL == 1 && !A_MN), "the weight-stationary path is single-CTA with a K-major A");
  constexpr int CDIV = CTA2 ? 2 : CL;                    // CTAs that share one tile index
  constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1u);
  const uint32_t rank = CDIV > 1 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator drained (every epilogue warp arrives)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* bfull_bar = tempty_bar + 3;              // weight-stationary: the resident B tile has landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fused column sums (bias gradients): accumulated per CTA in shared memory and flushed with ONE global atomic per
  // column at the end -- one atomic per column per 32-row chunk (766 k of them on 1024 addresses for the FFN dgrad) cost
  // more than the rest of the epilogue
  float* cs_smem = reinterpret_cast<float*>(smem + Cfg::CS_OFF);
  const int64_t cs_base = BSTAT ? static_cast<int64_t>(static_cast<int>(blockIdx.x) % (p.n_tiles > 0 ? p.n_tiles : 1)) * BN : 0;   // first column kept
  const int cs_n = BSTAT ? BN : static_cast<int>(p.N);
  const bool cs_in_smem = EXTRA && p.ep.colsum != nullptr && (BSTAT || p.N <= static_cast<int64_t>(Cfg::CS_MAXN));
  if (cs_in_smem)
    for (int i = threadIdx.x; i < cs_n; i += GEMM_THREADS) cs_smem[i] = 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);                      // multicast clusters: every CTA's MMAs release the stage
    }
    for (int s = 0; s < Cfg::AUX_BUFS * EPI_WARPS; ++s) mbar_init(reinterpret_cast<uint64_t*>(smem + Cfg::AUXBAR_OFF) + s, 1);
    mbar_init(bfull_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CTA2 ? 2 * EPI_WARPS : EPI_WARPS);     // pair: both CTAs' epilogues release the leader
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CDIV > 1) cluster_sync_all();            // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                    // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      TileCoord tc;
      if constexpr (BSTAT) {
        // the CTA's n-block of B, all k-blocks, once
        if (tile_at<BSTAT, CDIV>(p, 0, rank, tc)) {
          mbar_expect_tx(bfull_bar, static_cast<uint32_t>(p.total_kb) * Cfg::B_BYTES);
          for (int kb = 0; kb < p.total_kb; ++kb) {
            uint8_t* sBr = smem + Cfg::BRES_OFF + kb * Cfg::B_BYTES;
            if constexpr (!B_MN) {
              tma_load_2d(sBr, &tmB, bfull_bar, kb * BK, tc.n_blk * BN);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_2d(sBr + c * 8192, &tmB, bfull_bar, tc.n_blk * BN + c * 64, kb * BK);
            }
          }
        }
      }
      for (int pit = 0; tile_at<BSTAT, CDIV>(p, pit, rank, tc); ++pit) {
        const int n_blk = tc.n_blk, m_blk = tc.m_blk, split = tc.split;
        const bool half_tile = tc.hh >= 0;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kblocks_per_split);
        const int b_row0 = n_blk * BN + (CTA2 ? static_cast<int>(rank) * BNH : 0) + (tc.hh > 0 ? BN / 2 : 0);   // pair: this CTA's half of B's N rows
        for (int kb = kb0; kb < kb1; ++kb) {
          B2C_GT(0, pit, 2 * (kb - kb0));
          mbar_wait(&empty_bar[s], ph ^ 1u);
          B2C_GT(0, pit, 2 * (kb - kb0) + 1);
          uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          if constexpr (BSTAT) {
            mbar_expect_tx(&full_bar[s], Cfg::A_BYTES);
            tma_load_2d(sA, &tmA, &full_bar[s], kb * BK, m_blk * BM);
          } else if constexpr (!CTA2) {
            mbar_expect_tx(&full_bar[s], half_tile ? Cfg::A_BYTES + Cfg::B_BYTES / 2 : Cfg::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(sA, &tmA, &full_bar[s], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(sA + c * 8192, &tmA, &full_bar[s], m_blk * BM + c * 64, kb * BK);
            }
            if constexpr (CL == 1) {
              if constexpr (!B_MN) {
                if (half_tile) tma_load_2d(sB, &tmBh, &full_bar[s], kb * BK, b_row0);
                else tma_load_2d(sB, &tmB, &full_bar[s], kb * BK, b_row0);
              } else {
#pragma unroll
                for (int c = 0; c < BNH / 64; ++c)
                  if (!half_tile || c < BNH / 128) tma_load_2d(sB + c * 8192, &tmB, &full_bar[s], b_row0 + c * 64, kb * BK);
              }
            } else {
              // this CTA's 1/CL slice of the B tile, delivered to the same smem offset (and full barrier) of all CL CTAs
              if constexpr (!B_MN) {
                constexpr int RS = BN / CL;            // rows per slice (tensor-map box height)
                tma_load_2d_mc(sB + rank * (RS * 128), &tmB, &full_bar[s], kb * BK, b_row0 + static_cast<int>(rank) * RS, CL_MASK);
              } else {
                constexpr int CS = (BN / 64) / CL;     // 64-column chunks per slice
                static_assert(CS >= 1, "cluster too large for this tile width");
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                  const int cc = static_cast<int>(rank) * CS + c;
                  tma_load_2d_mc(sB + cc * 8192, &tmB, &full_bar[s], b_row0 + cc * 64, kb * BK, CL_MASK);
                }
              }
            }
          } else {
            // both CTAs' bytes land on the LEADER's full barrier (it alone waits before issuing the pair MMA)
            const uint32_t lead_bar = mapa_u32(&full_bar[s], 0);
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d_2sm(sA, &tmA, lead_bar, kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(sA + c * 8192, &tmA, lead_bar, m_blk * BM + c * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d_2sm(sB, &tmB, lead_bar, kb * BK, b_row0);
            } else {
#pragma unroll
              for (int c = 0; c < BNH / 64; ++c) tma_load_2d_2sm(sB + c * 8192, &tmB, lead_bar, b_row0 + c * 64, kb * BK);
            }
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0 && (!CTA2 || rank == 0)) {        // pairs: the leader issues for both CTAs
      constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 2 * BM : BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      constexpr uint32_t idesc_half = umma_idesc_bf16(CTA2 ? 2 * BM : BM, BN / 2, A_MN, B_MN);
      TileCoord tc;
      if constexpr (BSTAT) {
        if (tile_at<BSTAT, CDIV>(p, 0, rank, tc)) mbar_wait(bfull_bar, 0u);
      }
      for (uint32_t it = 0; tile_at<BSTAT, CDIV>(p, static_cast<int>(it), rank, tc); ++it) {
        const int split = tc.split;
        const uint32_t idesc_t = tc.hh >= 0 ? idesc_half : idesc;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(p.total_kb, kb0 + p.kblocks_per_split);
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        B2C_GT(1, it, 0);
        mbar_wait(&tempty_bar[as], aph ^ 1u);          // epilogue has drained this accumulator buffer
        B2C_GT(1, it, 1);
        tc_fence_after();
        const uint32_t tm_acc = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          B2C_GT(1, it, 2 + 2 * (kb - kb0));
          mbar_wait(&full_bar[s], ph);
          B2C_GT(1, it, 3 + 2 * (kb - kb0));
          tc_fence_after();
          const uint32_t a_addr = base + s * Cfg::STAGE_BYTES;
          const uint32_t b_addr = BSTAT ? base + Cfg::BRES_OFF + (kb - kb0) * Cfg::B_BYTES + (tc.hh > 0 ? Cfg::B_BYTES / 2 : 0u)
                                        : a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : umma_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : umma_desc_sw128(b_addr + k * 32, 16, 1024);
            if constexpr (CTA2) umma_bf16_2sm(tm_acc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tm_acc, da, db, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot free once these MMAs retire (pair: in both CTAs)
          if constexpr (CTA2) tc_commit_2sm(&empty_bar[s]);
          else if constexpr (CL > 1) tc_commit_mc(&empty_bar[s], CL_MASK);
          else tc_commit(&empty_bar[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
        }
        // accumulator complete -> epilogue (pair: each CTA's epilogue drains its own 128 rows)
        if constexpr (CTA2) tc_commit_2sm(&tfull_bar[as]); else tc_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ================================ epilogue (8 warps) ================================
    const b2c_epilogue& ep = p.ep;
    const DropCfg drop = drop_resolve(p.drop);
    constexpr int CW = Cfg::CW;
    constexpr int LPR = CW / 4;                        // lanes per row in the coalesced domain (float4 each)
    constexpr int RPI = 32 / LPR;                      // rows per warp instruction
    constexpr int NIT = 32 / RPI;                      // iterations to cover the warp's 32 rows
    constexpr int SWZ = LPR < 8 ? LPR - 1 : 7;         // XOR swizzle mask of the transpose buffer (chunks per row - 1, max 7)
    const int ew = warp - 2;
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int part = ew >> 2;                          // which slice of the BN columns
    float* stg = reinterpret_cast<float*>(smem + Cfg::STG_OFF + ew * Cfg::STG_WARP_BYTES);
    const int c_lane = lane % LPR;                     // float4 chunk (column group) this lane owns after the transpose
    const int r_lane = lane / LPR;
    const uint32_t tempty_lead0 = CTA2 ? mapa_u32(&tempty_bar[0], 0) : 0u;
    // SiLU-backward dgrad in the row domain (ROWAUX): the saved pre-activation arrives as 32 x 32 bf16 boxes by TMA, ONE
    // CHUNK AHEAD of its use, into the two 2 KB halves of the warp's staging buffer; the thread reads its own row of the
    // box, overwrites it in place with the output row, and the box leaves again through a TMA store.
    constexpr bool ROWAUX = EPI == EPI_SILU_BWD && Cfg::CW == 32;
    uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + Cfg::AUXBAR_OFF) + Cfg::AUX_BUFS * ew;
    uint32_t job = 0;                                  // chunks this warp has handled: buffer = job & 1
    [[maybe_unused]] uint32_t jbuf = 0;                // AUX3: job % 3
    uint32_t aux_ph = 0;                               // bit b = parity the next wait on aux_bar[b] uses
    constexpr int HALF_PARTS = (EPI_WARPS / 4) / 2;    // column slices that exist in a half-width tile
    // row-domain epilogues read the bias from a shared-memory copy (one broadcast LDS.128 per four columns instead of a
    // guarded global load whose L2 latency sat in front of every chunk's math); columns >= N read as 0 and are clipped
    // (the host only picks these epilogues when the padded width n_tiles * BN fits the copy)
    [[maybe_unused]] float* bias_s = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF);
    if constexpr ((EPI == EPI_BF16 || EPI == EPI_SILU) && Cfg::CW == 32) {
      const int ncols = BSTAT ? BN : p.n_tiles * BN;
      const float* bsrc = static_cast<const float*>(ep.bias);
      const int64_t c0 = BSTAT ? static_cast<int64_t>(static_cast<int>(blockIdx.x) % p.n_tiles) * BN : 0;
      for (int i = static_cast<int>(threadIdx.x) - 64; i < ncols; i += EPI_WARPS * 32)
        bias_s[i] = (bsrc && c0 + i < p.N) ? __ldg(bsrc + c0 + i) : 0.f;
      named_barrier_sync(1, EPI_WARPS * 32);
    }
    // lane 0: fetch the aux box of chunk ch_n (may run past the tile: it carries into the following tiles) of this CTA's
    // it_n-th tile into buffer jb & 1 -- or, L2ONLY, just pull it into L2 ahead of that fetch
    auto aux_issue = [&](int it_n, int ch_n, uint32_t buf, bool l2only) {
      it_n += ch_n / Cfg::CHUNKS;
      ch_n %= Cfg::CHUNKS;
      TileCoord t;
      if (!tile_at<BSTAT, CDIV>(p, it_n, rank, t)) return;
      if (t.hh >= 0 && part >= HALF_PARTS) return;
      const int64_t c = static_cast<int64_t>(t.n_blk) * BN + (t.hh > 0 ? BN / 2 : 0) + part * Cfg::SPAN + ch_n * Cfg::CW;
      const int64_t r = static_cast<int64_t>(t.m_blk) * BM + q * 32;
      if (c >= p.N || r >= p.M) return;                // the consumer skips the wait under the same condition
      if (l2only) {
        tma_prefetch_l2_2d(&tmO2, static_cast<int>(c), static_cast<int>(r));
        return;
      }
      uint64_t* b = aux_bar + buf;
      mbar_expect_tx(b, 2048);
      tma_load_2d(reinterpret_cast<uint8_t*>(stg) + buf * 2048, &tmO2, b, static_cast<int>(c), static_cast<int>(r));
    };
    if constexpr (ROWAUX) {
      if (lane == 0) {
        aux_issue(0, 0, 0, false);
        if constexpr (AUX3) aux_issue(0, 1, 1, false);
      }
    }
    TileCoord tc;
    for (uint32_t it = 0; tile_at<BSTAT, CDIV>(p, static_cast<int>(it), rank, tc); ++it) {
      const int m_blk = tc.m_blk, split = tc.split;
      const int64_t ncol0 = static_cast<int64_t>(tc.n_blk) * BN + (tc.hh > 0 ? BN / 2 : 0);   // first global column of the tile
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      if (tc.hh >= 0 && part >= HALF_PARTS) {          // half-width tail tile: this warp's columns do not exist
        mbar_wait(&tfull_bar[as], aph);                // (stay in step with the tile sequence before releasing the buffer)
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        continue;
      }
      const float* bias = split == 0 ? static_cast<const float*>(ep.bias) : nullptr;   // once per split-K
      constexpr bool G = EPI == EPI_GENERIC;
      const bool want_resid = G ? (ep.resid && !ep.accumulate && ep.out_dtype == B2C_F32) : (EPI == EPI_RESID);
      const bool want_aux = G ? (ep.act == B2C_ACT_SILU_BWD) : (EPI == EPI_SILU_BWD);
      const bool do_silu = G ? (ep.act == B2C_ACT_SILU) : (EPI == EPI_SILU);
      const bool out_bf16 = G ? (ep.out_dtype == B2C_BF16) : (EPI == EPI_BF16 || EPI == EPI_SILU || EPI == EPI_SILU_BWD);
      const bool do_accum = G ? (ep.accumulate != 0) : (EPI == EPI_ACCUM);
      const bool do_out2 = (G || (EXTRA && EPI == EPI_SILU)) && ep.out2 != nullptr;
      const bool do_drop = (G || (EXTRA && (EPI == EPI_SILU || EPI == EPI_RESID))) && p.drop.thresh16 != 0;
      const bool do_alpha = (G || EPI == EPI_RESID || EPI == EPI_ACCUM || EPI == EPI_SILU_BWD) && ep.alpha != 1.f;
      // ---- bf16 outputs without residual / aux operands: everything happens in the row-per-thread domain the TMEM
      // load delivers.  Bias / SiLU / dropout are elementwise, the 32 x 32 chunk is packed to bf16 and parked as a
      // 64B-swizzled 2 KB box (conflict-free 16-byte stores; the saved pre-activation gets a second box), and ONE lane
      // hands it to the TMA store unit, which also clips at M and N.  Against the fp32 transpose path this halves the
      // epilogue's shared-memory traffic (the resource these skinny-K GEMMs are bound by) and removes the LDS / address
      // arithmetic / STG stream entirely -- with ~28 KB of L1 left the STGs alone were a quarter of the epilogue.
      constexpr bool ROWDOM = (EPI == EPI_BF16 || EPI == EPI_SILU) && Cfg::CW == 32;
      if constexpr (ROWDOM) {
        const int64_t grow0 = static_cast<int64_t>(m_blk) * BM + q * 32;
        const uint64_t drow = static_cast<uint64_t>(grow0 + lane) * p.N;
        const uint32_t tm_row = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16) + part * Cfg::SPAN;
        const int colt0 = part * Cfg::SPAN;                                  // first column of this warp's slice inside the tile
        const float* bsp = bias_s + (BSTAT ? (tc.hh > 0 ? BN / 2 : 0) : ncol0) + colt0;   // bias of the slice (shared memory copy)
        // p = 0.5 (the reference's dropout, asrnn.py:29): ONE hash decides the thread's 32 consecutive columns, one bit
        // each.  This mask has no second consumer (the backward reads it off the saved pre-activation), so its
        // granularity and bit order are free; other probabilities use the 8-bit-per-element generator.
        const int mode = !do_drop ? 0 : (p.drop.thresh16 == 32768u ? 1 : 2);
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after();
        __syncwarp();
        // The chunk loop is software-pipelined: chunk ch+1's tcgen05.ld is in flight while chunk ch is computed, and the
        // bulk store of a chunk is only waited for (its shared-memory read) right before its box is rewritten -- with two
        // boxes per warp that is a whole chunk later.  Elementwise math runs on packed fp32 pairs.
        auto run = [&](auto MODE_) {
          constexpr int MODE = decltype(MODE_)::value;           // 0 no dropout, 1 one-bit masks (p = 0.5), 2 general p
          constexpr int SUBS = Cfg::SPAN / 16;                   // 16-column steps: two per 32-column store box
          uint32_t v[2][16];
          tmem_ld16(tm_row, v[0]);
          [[maybe_unused]] uint32_t keep_bits = 0;
          uint8_t* box = nullptr;
          const uint64_t half2 = pk2(0.5f, 0.5f);
          const uint32_t sw = (lane >> 1) & 3;                   // SWIZZLE_64B: 16-byte chunk index ^ ((row >> 1) & 3)
#pragma unroll
          for (int sc = 0; sc < SUBS; ++sc) {
            const int ch = sc >> 1, hf = sc & 1;
            uint32_t* vc = v[sc & 1];
            const int64_t col0 = ncol0 + colt0 + ch * CW;        // first column of the 32-column box this step belongs to
            tmem_wait_ld();
            if (sc + 1 < SUBS) {
              tmem_ld16(tm_row + (sc + 1) * 16, v[(sc + 1) & 1]);
            } else {                                             // all of this warp's TMEM reads for the tile are done
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_cluster(tempty_lead0 + as * 8); else mbar_arrive(&tempty_bar[as]);
              }
            }
            if (hf == 0) {
              if (ew == 0 && lane == 0) B2C_GT(2, it, 2 + 5 * ch);
              if constexpr (MODE == 1) keep_bits = drop_hash32(drop.seed ^ 0x2545F4914F6CDD1Dull, (drow + static_cast<uint64_t>(col0)) >> 5);
              // two 2 KB boxes per warp: with a single output they alternate (the store issued a chunk ago may still be
              // reading the other one); with the saved pre-activation both are rewritten every chunk
              box = reinterpret_cast<uint8_t*>(stg) + (do_out2 ? 0u : (job & 1u) * 2048u);
              if (do_out2) {                                     // (waiting here keeps the packed words out of local memory)
                if (lane == 0) tma_store_wait_read_n<0>();
                __syncwarp();
              }
            }
            uint32_t ph[8], pu[8];
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              const int g = 4 * hf + gg;                         // group of four columns inside the 32-column box
              const float4 bv = *reinterpret_cast<const float4*>(bsp + ch * CW + 4 * g);   // one address per warp: broadcast
              uint64_t z[2] = {add2(pk2u(vc[4 * gg], vc[4 * gg + 1]), pk2(bv.x, bv.y)),
                               add2(pk2u(vc[4 * gg + 2], vc[4 * gg + 3]), pk2(bv.z, bv.w))};
              if constexpr (MODE == 2) {                         // general p: per-element scales, scalar math
                float f[4], ds[4];
                upk2(z[0], f[0], f[1]);
                upk2(z[1], f[2], f[3]);
                drop_scale4g(drop, drow + static_cast<uint64_t>(col0 + 4 * g), ds[0], ds[1], ds[2], ds[3]);
                if (do_out2) {                                   // dropped positions carry B2C_DROPPED_PREACT (see below)
                  pu[2 * gg] = pack_bf16(ds[0] != 0.f ? f[0] : B2C_DROPPED_PREACT, ds[1] != 0.f ? f[1] : B2C_DROPPED_PREACT);
                  pu[2 * gg + 1] = pack_bf16(ds[2] != 0.f ? f[2] : B2C_DROPPED_PREACT, ds[3] != 0.f ? f[3] : B2C_DROPPED_PREACT);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = (do_silu ? f[j] * sigmoid_fast(f[j]) : f[j]) * ds[j];
                ph[2 * gg] = pack_bf16(f[0], f[1]);
                ph[2 * gg + 1] = pack_bf16(f[2], f[3]);
              } else {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const int j = 2 * g + h2;                      // pair index inside the box: columns 2j, 2j + 1
                  uint64_t r = z[h2];
                  if (do_silu) {                                 // z * sigmoid(z) = hz + hz * tanh(hz), hz = z / 2  (x 2 = 1/keep at p = 0.5)
                    const uint64_t hz = mul2(z[h2], half2);
                    float h0, h1;
                    upk2(hz, h0, h1);
                    const uint64_t t = pk2(tanh_approx(h0), tanh_approx(h1));
                    r = MODE == 1 ? fma2(z[h2], t, z[h2]) : fma2(hz, t, hz);
                  } else if (MODE == 1) {
                    r = add2(z[h2], z[h2]);
                  }
                  uint32_t o = cvt_bf16x2(r);
                  if constexpr (MODE == 1) {
                    // pair j reads its two keep bits at positions 8 apart: (7 - j, 15 - j) for j < 8, (8 + j, 16 + j) above
                    const uint32_t m = halfmask_from_bits_7_15(j < 8 ? keep_bits << j : keep_bits >> (j + 1));
                    o &= m;
                    if (do_out2) pu[2 * gg + h2] = (cvt_bf16x2(z[h2]) & m) | (0xC61CC61Cu & ~m);   // bf16(B2C_DROPPED_PREACT) where dropped
                  } else {
                    if (do_out2) pu[2 * gg + h2] = cvt_bf16x2(z[h2]);
                  }
                  ph[2 * gg + h2] = o;
                }
              }
            }
            if (hf == 0) {
              if (ew == 0 && lane == 0) B2C_GT(2, it, 3 + 5 * ch);
              if (!do_out2) {
                if (lane == 0) tma_store_wait_read_n<1>();
                __syncwarp();
              }
              if (ew == 0 && lane == 0) B2C_GT(2, it, 4 + 5 * ch);
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = 2 * hf + cc;
              *reinterpret_cast<uint4*>(box + lane * 64 + ((c ^ sw) << 4)) = make_uint4(ph[4 * cc], ph[4 * cc + 1], ph[4 * cc + 2], ph[4 * cc + 3]);
              if (do_out2)
                *reinterpret_cast<uint4*>(box + 2048 + lane * 64 + ((c ^ sw) << 4)) = make_uint4(pu[4 * cc], pu[4 * cc + 1], pu[4 * cc + 2], pu[4 * cc + 3]);
            }
            if (hf == 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0 && col0 < p.N && grow0 < p.M) {
                tma_store_2d(&tmO, box, static_cast<int>(col0), static_cast<int>(grow0));
                if (do_out2) tma_store_2d(&tmO2, box + 2048, static_cast<int>(col0), static_cast<int>(grow0));
              }
              if (lane == 0) tma_store_commit();                 // every chunk commits a (possibly empty) group: wait_group counts them
              ++job;
              if (ew == 0 && lane == 0) B2C_GT(2, it, 5 + 5 * ch);
            }
          }
        };
        using std::integral_constant;
        if (mode == 0) run(integral_constant<int, 0>{});
        else if (mode == 1) run(integral_constant<int, 1>{});
        else run(integral_constant<int, 2>{});
      }
      if constexpr (ROWAUX) {
        const int64_t grow0 = static_cast<int64_t>(m_blk) * BM + q * 32;
        const bool do_colsum = EXTRA && ep.colsum != nullptr;
        // SiLU'(u) = (1 + t + (u/2)(1 - t^2)) / 2, t = tanh(u/2); alpha and the 1/2 folded into ha.  Packed fp32 pairs.
        const uint64_t ha2 = pk2(0.5f * ep.alpha, 0.5f * ep.alpha), half2 = pk2(0.5f, 0.5f), mhalf2 = pk2(-0.5f, -0.5f), m1 = pk2(-1.f, -1.f);
#pragma unroll 1
        for (int ch = 0; ch < Cfg::CHUNKS; ++ch, ++job) {
          const int col_t = part * Cfg::SPAN + ch * CW;
          const int64_t col0 = ncol0 + col_t;
          const uint32_t b = AUX3 ? jbuf : (job & 1u);
          uint8_t* box = reinterpret_cast<uint8_t*>(stg) + b * 2048;
          const bool has = col0 < p.N && grow0 < p.M;
          if (ew == 0 && lane == 0) B2C_GT(2, it, 2 + 5 * ch);
          if constexpr (!AUX3) {
            if (lane == 0) {
              // two boxes: the other one's last store has read it: refill it with the next chunk's aux box (a whole chunk of
              // lead time; issuing it later -- after this chunk's math -- measured slower: 33.8 -> 37.7 us, the aux wait
              // moved into the open).  The wait itself is exposed: that store was issued moments ago (see AUX3 below).
              tma_store_wait_read_n<0>();
              aux_issue(static_cast<int>(it), ch + 1, (job + 1) & 1u, false);
            }
          }
          if (ch == 0) {
            if (ew == 0 && lane == 0) B2C_GT(2, it, 0);
            mbar_wait(&tfull_bar[as], aph);
            if (ew == 0 && lane == 0) B2C_GT(2, it, 1);
            tc_fence_after();
          }
          uint32_t v[32];
          __syncwarp();
          tmem_ld32(tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16) + col_t, v);
          const uint32_t sw = (lane >> 1) & 3;           // SWIZZLE_64B: 16-byte chunk index ^ ((row >> 1) & 3)
          uint4 a[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) a[c] = make_uint4(0u, 0u, 0u, 0u);
          if (has) {
            mbar_wait(aux_bar + b, (aux_ph >> b) & 1u);
            aux_ph ^= 1u << b;
#pragma unroll
            for (int c = 0; c < 4; ++c) a[c] = *reinterpret_cast<const uint4*>(box + lane * 64 + ((c ^ sw) << 4));
          }
          tmem_wait_ld();
          if (ew == 0 && lane == 0) B2C_GT(2, it, 3 + 5 * ch);
          if (ch == Cfg::CHUNKS - 1) {                   // all of this warp's TMEM reads for the tile are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTA2) mbar_arrive_cluster(tempty_lead0 + as * 8); else mbar_arrive(&tempty_bar[as]);
            }
          }
          uint64_t f2[16];                               // fp32 outputs as pairs (kept for the fused bias gradient)
          uint32_t o[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t w[4] = {a[c].x, a[c].y, a[c].z, a[c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t u2 = pk2u(w[j] << 16, w[j] & 0xFFFF0000u);          // bf16 pair -> fp32 pair
              const uint64_t hu = mul2(u2, half2), hn = mul2(u2, mhalf2);
              float h0, h1;
              upk2(hu, h0, h1);
              const uint64_t t = pk2(tanh_approx(h0), tanh_approx(h1));
              const uint64_t r = fma2(hn, fma2(t, t, m1), t);                    // t + (u/2)(1 - t^2)
              const int e = 4 * c + j;
              f2[e] = mul2(pk2u(v[2 * e], v[2 * e + 1]), fma2(ha2, r, ha2));
              o[e] = cvt_bf16x2(f2[e]);
            }
          }
          if (ew == 0 && lane == 0) B2C_GT(2, it, 4 + 5 * ch);
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(box + lane * 64 + ((c ^ sw) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (has) tma_store_2d(&tmO, box, static_cast<int>(col0), static_cast<int>(grow0));
            tma_store_commit();                            // (possibly empty: the wait below counts groups)
            if constexpr (AUX3) {
              // three boxes: the store issued a whole chunk ago has read its box -- no exposed wait -- and that box takes the
              // aux operand of the chunk after next, which then has a whole chunk to arrive
              tma_store_wait_read_n<1>();
              aux_issue(static_cast<int>(it), ch + 2, jbuf == 0 ? 2u : jbuf - 1u, false);
            }
          }
          if constexpr (AUX3) jbuf = jbuf == 2 ? 0u : jbuf + 1u;
          if (ew == 0 && lane == 0) B2C_GT(2, it, 5 + 5 * ch);
          if (do_colsum) {
            // bias gradient from the fp32 values: transpose-reduce over the warp's 32 rows -- 31 shuffles leave lane L
            // with the sum of column col0 + L (each step folds the half of the columns whose bit differs from the lane's)
            float f[32];
#pragma unroll
            for (int e = 0; e < 16; ++e) upk2(f2[e], f[2 * e], f[2 * e + 1]);
            float r16[16], r8[8], r4[4], r2[2];
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              r16[i] = (b16 ? f[i + 16] : f[i]) + __shfl_xor_sync(0xffffffffu, b16 ? f[i] : f[i + 16], 16);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              r8[i] = (b8 ? r16[i + 8] : r16[i]) + __shfl_xor_sync(0xffffffffu, b8 ? r16[i] : r16[i + 8], 8);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              r4[i] = (b4 ? r8[i + 4] : r8[i]) + __shfl_xor_sync(0xffffffffu, b4 ? r8[i] : r8[i + 4], 4);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              r2[i] = (b2 ? r4[i + 2] : r4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? r4[i] : r4[i + 2], 2);
            const float tot = (b1 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? r2[0] : r2[1], 1);
            if (col0 + lane < p.N) {
              if (cs_in_smem) atomicAdd(cs_smem + (col0 - cs_base) + lane, tot); else atomicAdd(ep.colsum + col0 + lane, tot);
            }
          }
        }
      }
      if constexpr (!ROWDOM && !ROWAUX)
#pragma unroll 1
      for (int ch = 0; ch < Cfg::CHUNKS; ++ch) {
        const int col_t = part * Cfg::SPAN + ch * CW;  // first column of this chunk inside the tile
        const int64_t col = ncol0 + col_t + c_lane * 4;
        const bool col_ok = col < p.N;
        const int64_t row0 = static_cast<int64_t>(m_blk) * BM + q * 32 + r_lane;
        const int64_t rl64 = p.M - row0;
        const int rows_left = rl64 > 32 ? 32 : static_cast<int>(rl64);      // rows row0 + i*RPI are valid while i*RPI < rows_left
        // touch every line this chunk will READ (residual / saved pre-activation) BEFORE waiting for the accumulator:
        // prefetch.global.L2 needs no registers, and the later loads then see L2-hit latency that the 4-way
        // unrolled loop below covers
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((G || EPI != EPI_ACCUM) && bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
        const float* rp = static_cast<const float*>(ep.resid) + row0 * ep.ld_out + col;
        const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(ep.aux) + row0 * ep.ld_aux + col;
        // every global READ of this chunk (residual / saved pre-activation) is issued into registers BEFORE waiting
        // for the accumulator, so its DRAM/L2 latency overlaps this tile's main loop (NIT <= 8: at most 32 registers)
        // (separate register arrays per stream and loads straight into them: routing the 8-byte aux load through a
        // uint4 slot made the compiler emit a MOV behind every LDG, which blocks in-order issue until that load returns
        // -- the eight "prefetches" ran one memory latency after the other, 3 k cycles per tile)
        if (ew == 0 && lane == 0) B2C_GT(2, it, 8 + 2 * ch);
        uint4 pre[NIT];
        uint2 prea[NIT];
        if (col_ok && want_resid) {
#pragma unroll
          for (int i = 0; i < NIT; ++i)
            if (i * RPI < rows_left) pre[i] = ldg_stream_v4(rp + static_cast<int64_t>(i) * RPI * ep.ld_out);
        } else if (col_ok && want_aux) {
#pragma unroll
          for (int i = 0; i < NIT; ++i)
            if (i * RPI < rows_left) prea[i] = ldg_stream_v2(ap + static_cast<int64_t>(i) * RPI * ep.ld_aux);
        }
        if (ew == 0 && lane == 0) B2C_GT(2, it, 9 + 2 * ch);
        if (ch == 0) {
          if (ew == 0 && lane == 0) B2C_GT(2, it, 0);
          mbar_wait(&tfull_bar[as], aph);
          if (ew == 0 && lane == 0) B2C_GT(2, it, 1);
          tc_fence_after();
        }
        uint32_t v[CW];
        __syncwarp();
        const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16) + col_t;
        if constexpr (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tmem_wait_ld();
        if (ew == 0 && lane == 0) B2C_GT(2, it, 2 + 2 * ch);
        if (ch == Cfg::CHUNKS - 1) {                   // all of this warp's TMEM reads for the tile are done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_cluster(tempty_lead0 + as * 8); else mbar_arrive(&tempty_bar[as]);
          }
        }
        // transpose through smem: row = lane, 16-byte chunk c stored at (c ^ (lane & SWZ))
#pragma unroll
        for (int c = 0; c < LPR; ++c)
          *reinterpret_cast<uint4*>(stg + lane * CW + ((c ^ (lane & SWZ)) << 2)) =
              make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        const bool do_colsum = (G || (EXTRA && EPI == EPI_SILU_BWD)) && ep.colsum != nullptr && out_bf16;
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
        if (col_ok) {
          // loop-carried pointers / indices: one 64-bit add per stream per row step instead of re-deriving
          // base + i*stride (IMAD.WIDE chains) in every unrolled iteration -- the epilogue is issue-bound
          uint64_t di = static_cast<uint64_t>(row0) * p.N + col;            // dropout element index (even)
          const uint64_t di_step = static_cast<uint64_t>(RPI) * p.N;
          __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(ep.out) + row0 * ep.ld_out + col;
          float* of = static_cast<float*>(ep.out) + row0 * ep.ld_out + col;
          __nv_bfloat16* o2 = static_cast<__nv_bfloat16*>(ep.out2) + row0 * ep.ld_out2 + col;
          const int64_t o_step = static_cast<int64_t>(RPI) * ep.ld_out, o2_step = static_cast<int64_t>(RPI) * ep.ld_out2;
          // transpose-buffer read offsets: (r & SWZ) cycles with period PC rows steps, so PC precomputed offsets suffice
          constexpr int PC = (SWZ + 1) / RPI > 0 ? (SWZ + 1) / RPI : 1;
          int rd_off[PC];
#pragma unroll
          for (int pc = 0; pc < PC; ++pc) {
            const int r = pc * RPI + r_lane;
            rd_off[pc] = r * CW + ((c_lane ^ (r & SWZ)) << 2);
          }
          // variants that index the prefetch registers need the full unroll; the others stay compact (I-cache)
          constexpr int UNR = (EPI == EPI_RESID || EPI == EPI_SILU_BWD || EPI == EPI_GENERIC) ? NIT : (NIT < 4 ? NIT : 4);
#pragma unroll UNR
          for (int i = 0; i < NIT; ++i) {
            if (i * RPI >= rows_left) break;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + rd_off[i % PC] + (i / PC) * (PC * RPI * CW));
            float f[4] = {a4.x + bv.x, a4.y + bv.y, a4.z + bv.z, a4.w + bv.w};
            float ds0 = 1.f, ds1 = 1.f, ds2 = 1.f, ds3 = 1.f;     // dropout scales of this float4 (1 = dropout off)
            if (do_drop) {
              drop_scale4g(drop, di, ds0, ds1, ds2, ds3);
              di += di_step;
            }
            if (do_out2) {
              // pre-activation copy (saved for the backward pass).  Where the activation is about to be DROPPED the copy
              // holds B2C_DROPPED_PREACT instead: SiLU'(-1e4) is exactly 0, so the backward dgrad (act = SILU_BWD,
              // alpha = 1/keep) needs no mask regeneration -- the dropout hash costs a third of that kernel's epilogue
              const float u0 = ds0 != 0.f ? f[0] : B2C_DROPPED_PREACT, u1 = ds1 != 0.f ? f[1] : B2C_DROPPED_PREACT;
              const float u2 = ds2 != 0.f ? f[2] : B2C_DROPPED_PREACT, u3 = ds3 != 0.f ? f[3] : B2C_DROPPED_PREACT;
              *reinterpret_cast<uint2*>(o2) = make_uint2(pack_bf16(u0, u1), pack_bf16(u2, u3));
              o2 += o2_step;
            }
            if (do_silu) {
#pragma unroll
              for (int j = 0; j < 4; ++j) f[j] *= sigmoid_fast(f[j]);
            } else if (want_aux) {                      // f = dL/dh ; aux = pre-activation u ; out = f * silu'(u)
              const float2 a0 = unpack_bf16(prea[i].x), a1 = unpack_bf16(prea[i].y);
              const float u[4] = {a0.x, a0.y, a1.x, a1.y};
              // SiLU'(u) = s + u s (1 - s), s = sigmoid(u) = (1 + t)/2, t = tanh(u/2)  =>  SiLU' = (1 + t + (u/2)(1 - t^2)) / 2:
              // one MUFU and three FFMAs, with the 1/keep factor (alpha) folded into the last one
              const float ha = (G || EPI == EPI_SILU_BWD) ? 0.5f * ep.alpha : 0.5f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float hu = 0.5f * u[j];
                const float t = tanh_approx(hu);
                const float r = fmaf(hu, fmaf(-t, t, 1.f), t);
                f[j] *= fmaf(ha, r, ha);
              }
            }
            if (do_drop) { f[0] *= ds0; f[1] *= ds1; f[2] *= ds2; f[3] *= ds3; }
            if (do_alpha && !want_aux) {
#pragma unroll
              for (int j = 0; j < 4; ++j) f[j] *= ep.alpha;
            }
            if (out_bf16) {
              const uint2 packed = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
              *reinterpret_cast<uint2*>(ob) = packed;
              ob += o_step;
              if (do_colsum) { cs0 += f[0]; cs1 += f[1]; cs2 += f[2]; cs3 += f[3]; }   // bias gradient, from the fp32 values
            } else {
              if (do_accum) {                          // split-K / gradient accumulation: fp32 reductions in L2
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(of), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3])
                             : "memory");
              } else {
                if (want_resid) {
                  f[0] += __uint_as_float(pre[i].x); f[1] += __uint_as_float(pre[i].y);
                  f[2] += __uint_as_float(pre[i].z); f[3] += __uint_as_float(pre[i].w);
                }
                *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
              }
              of += o_step;
            }
          }
        }
        if (ew == 0 && lane == 0) B2C_GT(2, it, 3 + 2 * ch);
        if (do_colsum) {                               // warp-level: lanes sharing c_lane hold partials of the same columns
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, o); cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
            cs2 += __shfl_xor_sync(0xffffffffu, cs2, o); cs3 += __shfl_xor_sync(0xffffffffu, cs3, o);
          }
          if (r_lane == 0 && col_ok) {
            float* dst = cs_in_smem ? cs_smem + (col - cs_base) : ep.colsum + col;
            atomicAdd(dst, cs0); atomicAdd(dst + 1, cs1); atomicAdd(dst + 2, cs2); atomicAdd(dst + 3, cs3);
          }
        }
      }
    }
  }
  if (warp >= 2 && lane == 0) tma_store_wait_read();    // row-domain epilogues: shared memory outlives the last bulk stores
  tc_fence_before();
  __syncthreads();
  if (cs_in_smem)
    for (int i = threadIdx.x; i < cs_n; i += GEMM_THREADS) {
      const float v = cs_smem[i];
      if (v != 0.f && cs_base + i < p.N) atomicAdd(p.ep.colsum + cs_base + i, v);
    }
  if constexpr (CDIV > 1) cluster_sync_all();            // peers may still be multicasting into this CTA / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool CTA2, bool EXTRA, int CL = 1, bool BSTAT = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                       const CUtensorMap& tbh, GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTA2, BSTAT, (BSTAT && BN == 128 && EPI == EPI_SILU_BWD)>;
  constexpr int GEMM_THREADS = Cfg::THREADS;
  static_assert((EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM) <= 227 * 1024, "shared-memory plan does not fit");
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EPI, CTA2, EXTRA, CL, BSTAT>;
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM));
    configured = true;
  }
  p.n_tiles = static_cast<int>((p.N + BN - 1) / BN);
  constexpr int CDIV = CTA2 ? 2 : CL;
  p.m_tiles = static_cast<int>((p.M + CDIV * BM - 1) / (CDIV * BM));      // M units: blocks, pairs or clusters
  p.total_tiles = p.m_tiles * p.n_tiles * p.splits;
  p.tiles_full = p.total_tiles;
  p.bs_group = p.bs_rounds = p.bs_rem = p.bs_half = 0;
  if constexpr (BSTAT) {
    int group = num_sms() / p.n_tiles;
    if (group > p.m_tiles) group = p.m_tiles;
    p.bs_group = group;
    p.bs_rounds = p.m_tiles / group;
    p.bs_rem = p.m_tiles % group;
    p.bs_half = (p.bs_rem > 0 && 2 * p.bs_rem <= group) ? 1 : 0;
    B2C_CUDA(launch_k(kern, dim3(group * p.n_tiles), dim3(GEMM_THREADS), EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM, st, ta, tb, to, to2, tbh, p));
    B2C_CUDA(cudaGetLastError());
    return 0;
  }
  if constexpr (!CTA2 && CL == 1 && !A_MN && BN >= 128) {
    // a last wave that is at most half full is issued as half-width tiles (see decode_tile)
    static const bool tail_ok = [] { const char* e = getenv("B2C_GEMM_TAIL"); return !(e && e[0] == '0'); }();
    const int nsm = num_sms(), rem = p.total_tiles % nsm;
    if (tail_ok && p.splits == 1 && p.total_tiles > nsm && rem > 0 && 2 * rem <= nsm) {
      p.tiles_full = p.total_tiles - rem;
      p.total_tiles = p.tiles_full + 2 * rem;
    }
  }
  if constexpr (CL > 1) {
    static int max_clusters = -1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters < 0) {                     // co-resident clusters of this shape (GPC boundaries cost a few SMs)
      cfg.gridDim = dim3(CL * (num_sms() / CL));
      int n = 0;
      B2C_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      max_clusters = n > 0 ? n : num_sms() / CL;
    }
    const int ncl = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
    cfg.gridDim = dim3(CL * ncl);
    B2C_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, to2, tbh, p));
  } else if constexpr (!CTA2) {
    const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    B2C_CUDA(launch_k(kern, dim3(grid), dim3(GEMM_THREADS), EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM, st, ta, tb, to, to2, tbh, p));
  } else {
    const int pairs = num_sms() / 2;
    const int npair = p.total_tiles < pairs ? p.total_tiles : pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * npair);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = EXTRA ? Cfg::SMEM_EXTRA : Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2C_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, to2, tbh, p));
  }
  B2C_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2c

using namespace b2c;

#ifdef B2C_GEMM_TIMING
extern "C" int b2c_debug_gemm_timing(long long* host_out) {
  B2C_CUDA(cudaDeviceSynchronize());
  B2C_CUDA(cudaMemcpyFromSymbol(host_out, g_gemm_dbg, sizeof(long long) * 3 * 16 * 32));
  return 0;
}
#endif

extern "C" int b2c_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                             int64_t N, int64_t K, int split_k, int block_n, const b2c_epilogue* ep, void* stream) {
  B2C_REQUIRE(A && B && ep && ep->out, "b2c_gemm_bf16: null pointer");
  B2C_REQUIRE(M > 0 && N > 0 && K > 0, "b2c_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N,
              (long long)K);
  B2C_REQUIRE(N % 8 == 0, "b2c_gemm_bf16: N=%lld must be a multiple of 8", (long long)N);
  B2C_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "b2c_gemm_bf16: leading dims must be multiples of 8 elements (16 B)");
  B2C_REQUIRE(ep->ld_out % 4 == 0 && ep->ld_out2 % 4 == 0 && ep->ld_aux % 4 == 0,
              "b2c_gemm_bf16: output / aux pitches must be multiples of 4 elements");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0,
              "b2c_gemm_bf16: pointers must be 16-byte aligned");
  B2C_REQUIRE(!(ep->accumulate && ep->out_dtype != B2C_F32), "b2c_gemm_bf16: accumulate needs fp32 output");
  B2C_REQUIRE(!(ep->act == B2C_ACT_SILU_BWD && !ep->aux), "b2c_gemm_bf16: SILU_BWD needs aux");
  // block_n: 0 = automatic; 64/128/256 = one CTA per 128 x block_n tile; -128/-256 = CTA pair per 256 x |block_n| tile
  int bn = block_n;
  bool cta2 = false;
  int cl = 1;
  int bstat = 0;                                 // 1 = forced, 2 = picked automatically (falls back when no kernel matches)
  if (bn == 0) {
    // measured on B200 at the Conformer-M shapes (tools/kernel_bench.py, tools/wgrad_probe.py): weight gradients and
    // wide forward GEMMs run best on 128x256 tiles of single CTAs (weight gradients with a split-K that gives ONE wave:
    // 256x1024 at split 18 takes 16.8 us against 20.5 us for CTA pairs at split 37), the rest on 128x128.
    // B2C_GEMM_2SM=1 forces the CTA-pair path for weight gradients.
    static const bool force2 = [] { const char* e = getenv("B2C_GEMM_2SM"); return e && e[0] == '1'; }();
    if (a_mn && b_mn && N >= 256 && M >= 256) { bn = 256; cta2 = force2; }
    else if (ep->act == B2C_ACT_SILU_BWD) bn = N >= 512 ? 256 : ((N >= 128) ? 128 : 64);   // row-domain aux epilogue: 27.3 vs 28.9 us
    else if (N >= 512 || (N >= 256 && K >= 512)) bn = 256;
    else bn = (N >= 128) ? 128 : 64;
  } else if (bn < 0) {
    bn = -bn;
    cta2 = true;
    B2C_REQUIRE(bn == 128 || bn == 256, "b2c_gemm_bf16: CTA-pair tiles need block_n -128 or -256");
  } else if (bn >= 9000) {                       // 9000 + BN: force the weight-stationary path (tools / tests)
    bn -= 9000;
    bstat = 1;
    B2C_REQUIRE((bn == 128 || bn == 256) && !a_mn && K <= BS_MAXKB * BK && split_k <= 1,
                "b2c_gemm_bf16: the weight-stationary path needs block_n = 9000 + {128,256}, a K-major A and K <= %d", BS_MAXKB * BK);
  } else if (bn >= 1000) {                       // 1000*CL + BN: multicast cluster of CL CTAs along M (CL = 2, 4)
    cl = bn / 1000;
    bn %= 1000;
    B2C_REQUIRE((cl == 2 || cl == 4) && (bn == 128 || bn == 256) && !a_mn && !(b_mn && bn == 128 && cl == 4),
                "b2c_gemm_bf16: multicast clusters need block_n = 1000*{2,4} + {128,256} and a K-major A");
  }
  if (block_n == 0 && !cta2 && !a_mn && bn >= 128) {
    // automatic: share the B tile across vertically adjacent tiles when the problem is tall enough to fill the clusters
    static const int auto_cl = [] { const char* e = getenv("B2C_GEMM_CLUSTER"); return e ? atoi(e) : B2C_DEFAULT_CLUSTER; }();
    if ((auto_cl == 2 || auto_cl == 4) && M >= 128ll * auto_cl * 8) cl = (b_mn && bn == 128 && auto_cl == 4) ? 2 : auto_cl;
  }
  const int bn_plain = bn;                       // what the automatic choice falls back to when no stationary kernel matches
  if (block_n == 0 && !cta2 && cl == 1 && !a_mn && K <= BS_MAXKB * BK && split_k <= 1 && M >= 4 * BM && ep->out_dtype == B2C_BF16 &&
      !ep->resid && !ep->accumulate) {
    // K <= 256: a tile's main loop is at most 4 k-blocks and re-fetching the B tile for every 128 rows of A is two thirds of
    // what the tile pulls out of L2 (128 of 192 KB per 128x256 tile).  Where it measured faster (tools/bstat_probe.py, M =
    // 23872, K = 256: QKV 12.5 -> 11.2 us, pointwise-1 10.2 -> 9.4, SiLU-backward dgrad 33.8 -> 31.6, N = 256 dgrad
    // 8.0 -> 7.4; NOT the FFN1 SiLU epilogues: 24.0 streaming vs 25.7) the n-block's weights stay parked in shared memory.
    // B2C_GEMM_BSTAT=0: off.
    static const bool bs_auto = [] { const char* e = getenv("B2C_GEMM_BSTAT"); return !(e && e[0] == '0'); }();
    const bool extra_req = ep->out2 || ep->colsum || ep->drop_p > 0.f;
    int pick = 0;
    if (ep->act == B2C_ACT_NONE && !extra_req && !ep->aux && ep->alpha == 1.f && N % 256 == 0) pick = 256;
    else if (ep->act == B2C_ACT_SILU_BWD && N % 128 == 0) pick = 128;
    if (bs_auto && pick) {
      bstat = 2;
      bn = pick;
    }
  }
  B2C_REQUIRE(bn == 64 || bn == 128 || bn == 256, "b2c_gemm_bf16: block_n must be 64, 128 or 256");
  const int total_kb = static_cast<int>((K + BK - 1) / BK);
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > total_kb) splits = total_kb;
  B2C_REQUIRE(splits == 1 || ep->accumulate, "b2c_gemm_bf16: split_k > 1 needs accumulate=1 (fp32 atomics)");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.total_kb = total_kb;
  p.kblocks_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.m_tiles = static_cast<int>((M + BM - 1) / BM);
  p.ep = *ep;
  p.drop = make_drop(ep->drop_p, ep->drop_seed);
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = get_tensor_map_2d(&ta, A, 2, K, M, lda * 2, BK, BM, true);
  else       rc = get_tensor_map_2d(&ta, A, 2, M, K, lda * 2, 64, BK, true);
  if (rc) return rc;
  if (!b_mn) rc = get_tensor_map_2d(&tb, B, 2, K, N, ldb * 2, BK, cta2 ? bn / 2 : bn / cl, true);
  else       rc = get_tensor_map_2d(&tb, B, 2, N, K, ldb * 2, 64, BK, true);
  if (rc) return rc;
  CUtensorMap tbh = tb;                  // B boxes of half-width tail tiles (K-major B only; MN-major B loads fewer chunks)
  if (!b_mn && !cta2 && cl == 1 && bn >= 128 && (rc = get_tensor_map_2d(&tbh, B, 2, K, N, ldb * 2, BK, bn / 2, true))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // bf16 outputs of the plain / SiLU epilogues leave through TMA stores (row-domain epilogue): 32 x 32 boxes, 64B swizzle
  const bool rowdom_ok = ep->out_dtype == B2C_BF16 && bn >= 128 && ep->ld_out % 8 == 0 && (N + bn - 1) / bn * bn <= 2048 &&
                         (!ep->out2 || (ep->ld_out2 % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->out2) & 15) == 0));
  // ... and the SiLU-backward dgrad also READS its aux operand (the saved pre-activation) as such boxes
  const bool rowaux_ok = rowdom_ok && ep->act == B2C_ACT_SILU_BWD && !ep->out2 && ep->ld_aux % 8 == 0 &&
                         (reinterpret_cast<uintptr_t>(ep->aux) & 15) == 0;
  CUtensorMap to = ta, to2 = ta;
  if (rowdom_ok) {
    if ((rc = get_tensor_map_2d_sw(&to, ep->out, 2, N, M, ep->ld_out * 2, 32, 32, 2))) return rc;
    if (ep->out2 && (rc = get_tensor_map_2d_sw(&to2, ep->out2, 2, N, M, ep->ld_out2 * 2, 32, 32, 2))) return rc;
    if (rowaux_ok && (rc = get_tensor_map_2d_sw(&to2, ep->aux, 2, N, M, ep->ld_aux * 2, 32, 32, 2))) return rc;
  }
  // pick the specialised epilogue when the request matches one of the hot-path patterns, else the generic kernel
  int epi = EPI_GENERIC;
  const bool plain = !ep->out2 && !ep->aux && !ep->resid && !ep->accumulate && ep->act == B2C_ACT_NONE && ep->alpha == 1.f;
  if (ep->out_dtype == B2C_BF16 && plain && ep->drop_p == 0.f && !ep->colsum && (rowdom_ok || bn < 128)) epi = EPI_BF16;
  else if (ep->out_dtype == B2C_BF16 && ep->act == B2C_ACT_SILU && !ep->resid && ep->alpha == 1.f && !ep->colsum && !ep->aux &&
           (rowdom_ok || bn < 128)) epi = EPI_SILU;
  else if (ep->out_dtype == B2C_F32 && ep->resid && !ep->accumulate && ep->act == B2C_ACT_NONE && !ep->out2) epi = EPI_RESID;
  else if (ep->out_dtype == B2C_BF16 && ep->act == B2C_ACT_SILU_BWD && !ep->out2 && !ep->bias && ep->drop_p == 0.f &&
           (rowaux_ok || bn < 128)) epi = EPI_SILU_BWD;
  else if (ep->out_dtype == B2C_F32 && ep->accumulate && ep->act == B2C_ACT_NONE && !ep->out2 && !ep->bias && ep->drop_p == 0.f) epi = EPI_ACCUM;
  const bool extra = ep->out2 || ep->colsum || ep->drop_p > 0.f;
#define B2C_LAUNCH(BN_, AM, BMN, E, C2)                                                           \
  do {                                                                                             \
    if (extra) return launch_gemm<BN_, AM, BMN, E, C2, true>(ta, tb, to, to2, tbh, p, st);                       \
    return launch_gemm<BN_, AM, BMN, E, C2, false>(ta, tb, to, to2, tbh, p, st);                                 \
  } while (0)
#define B2C_LAUNCH1(BN_, AM, BMN, E, C2) return launch_gemm<BN_, AM, BMN, E, C2, true>(ta, tb, to, to2, tbh, p, st)
#define B2C_DISPATCH(BN_, C2)                                                                      \
  if (bn == BN_ && cta2 == C2) {                                                                   \
    if (!a_mn && !b_mn) {                                                                          \
      if (epi == EPI_BF16) B2C_LAUNCH(BN_, false, false, EPI_BF16, C2);                            \
      if (epi == EPI_SILU) B2C_LAUNCH(BN_, false, false, EPI_SILU, C2);                            \
      if (epi == EPI_RESID) B2C_LAUNCH(BN_, false, false, EPI_RESID, C2);                          \
      B2C_LAUNCH1(BN_, false, false, EPI_GENERIC, C2);                                             \
    }                                                                                              \
    if (!a_mn && b_mn) {                                                                           \
      if (epi == EPI_BF16) B2C_LAUNCH(BN_, false, true, EPI_BF16, C2);                             \
      if (epi == EPI_SILU_BWD) B2C_LAUNCH(BN_, false, true, EPI_SILU_BWD, C2);                     \
      B2C_LAUNCH1(BN_, false, true, EPI_GENERIC, C2);                                              \
    }                                                                                              \
    if (a_mn && b_mn) {                                                                            \
      if (epi == EPI_ACCUM) B2C_LAUNCH1(BN_, true, true, EPI_ACCUM, C2);                           \
      B2C_LAUNCH1(BN_, true, true, EPI_GENERIC, C2);                                               \
    }                                                                                              \
    B2C_LAUNCH1(BN_, true, false, EPI_GENERIC, C2);                                                \
  }
#define B2C_LAUNCH_CL(BN_, BMN, E, CL_)                                                            \
  do {                                                                                             \
    if (extra) return launch_gemm<BN_, false, BMN, E, false, true, CL_>(ta, tb, to, to2, tbh, p, st);            \
    return launch_gemm<BN_, false, BMN, E, false, false, CL_>(ta, tb, to, to2, tbh, p, st);                      \
  } while (0)
#define B2C_DISPATCH_CL(BN_, CL_, DG)                                                              \
  if (bn == BN_ && cl == CL_ && !cta2 && !a_mn) {                                                  \
    if (!b_mn) {                                                                                   \
      if (epi == EPI_BF16) B2C_LAUNCH_CL(BN_, false, EPI_BF16, CL_);                               \
      if (epi == EPI_SILU) B2C_LAUNCH_CL(BN_, false, EPI_SILU, CL_);                               \
      if (epi == EPI_RESID) B2C_LAUNCH_CL(BN_, false, EPI_RESID, CL_);                             \
      return launch_gemm<BN_, false, false, EPI_GENERIC, false, true, CL_>(ta, tb, to, to2, tbh, p, st);         \
    }                                                                                              \
    if constexpr (DG) {                                                                            \
      if (epi == EPI_BF16) B2C_LAUNCH_CL(BN_, true, EPI_BF16, CL_);                                \
      if (epi == EPI_SILU_BWD) B2C_LAUNCH_CL(BN_, true, EPI_SILU_BWD, CL_);                        \
      return launch_gemm<BN_, false, true, EPI_GENERIC, false, true, CL_>(ta, tb, to, to2, tbh, p, st);          \
    }                                                                                              \
  }
#define B2C_LAUNCH_BS(BN_, BMN, E)                                                                 \
  do {                                                                                             \
    if (extra) return launch_gemm<BN_, false, BMN, E, false, true, 1, true>(ta, tb, to, to2, tbh, p, st);        \
    return launch_gemm<BN_, false, BMN, E, false, false, 1, true>(ta, tb, to, to2, tbh, p, st);                  \
  } while (0)
#define B2C_DISPATCH_BS(BN_)                                                                       \
  if (bstat && bn == BN_) {                                                                        \
    if (!b_mn) {                                                                                   \
      if (epi == EPI_BF16) B2C_LAUNCH_BS(BN_, false, EPI_BF16);                                    \
      if (epi == EPI_SILU) B2C_LAUNCH_BS(BN_, false, EPI_SILU);                                    \
    } else {                                                                                       \
      if (epi == EPI_BF16) B2C_LAUNCH_BS(BN_, true, EPI_BF16);                                     \
      if (epi == EPI_SILU_BWD) B2C_LAUNCH_BS(BN_, true, EPI_SILU_BWD);                             \
    }                                                                                              \
  }
  B2C_DISPATCH_BS(128)
  B2C_DISPATCH_BS(256)
#undef B2C_DISPATCH_BS
#undef B2C_LAUNCH_BS
  B2C_REQUIRE(bstat != 1, "b2c_gemm_bf16: no weight-stationary kernel for this epilogue");
  if (bstat == 2 && bn != bn_plain) {            // generic epilogue: back to the default tile and its B boxes
    bn = bn_plain;
    if (!b_mn && (rc = get_tensor_map_2d(&tb, B, 2, K, N, ldb * 2, BK, bn / cl, true))) return rc;
    tbh = tb;
    if (!b_mn && bn >= 128 && (rc = get_tensor_map_2d(&tbh, B, 2, K, N, ldb * 2, BK, bn / 2, true))) return rc;
  }
  B2C_DISPATCH_CL(128, 2, true)
  B2C_DISPATCH_CL(128, 4, false)
  B2C_DISPATCH_CL(256, 2, true)
  B2C_DISPATCH_CL(256, 4, true)
#undef B2C_DISPATCH_CL
#undef B2C_LAUNCH_CL
  B2C_REQUIRE(cl == 1, "b2c_gemm_bf16: no multicast-cluster kernel for this operand layout / tile");
  B2C_DISPATCH(64, false)
  B2C_DISPATCH(128, false)
  B2C_DISPATCH(256, false)
  B2C_DISPATCH(128, true)
  B2C_DISPATCH(256, true)
#undef B2C_LAUNCH1
#undef B2C_LAUNCH
#undef B2C_DISPATCH
  return -1;
}
