// Flash-style multi-head self-attention on tcgen05 / TMEM / TMA (sm_100a), head dim 64 (smaller heads are
// zero-padded to 64 by the host side).
//
// Replaces scaled_dot_product_attention reached from nn.MultiheadAttention (torch/nn/functional.py:6608-6682),
// called by torchaudio ConformerLayer (conformer.py:152,194-200): softmax(Q K^T / sqrt(dh) + key mask) V where
// keys t >= lengths[b] are masked and padded QUERY rows are still computed.
//
// Forward, one CTA = (128-query tile, head, utterance), 256 threads, two threads per query row / TMEM lane:
//   TMA   Q tile once, then per 128-key block K_j, V_j                       (128B-swizzled smem)
//   UMMA  S = Q K_j^T          (M128 N128 K16 x4, both operands K-major)      -> TMEM cols [0,128)
//   regs  online softmax on the thread's row (tcgen05.ld), P -> smem as bf16 in the K-major SW128 UMMA layout
//   UMMA  O_j = P V_j          (M128 N64 K16 x8, A K-major, B MN-major = V as loaded) -> TMEM cols [128,192)
//   regs  o = o*alpha + O_j    (running output stays in registers; no TMEM rescale pass)
// Key blocks past lengths[b] are skipped entirely.  80 KB smem and 256 TMEM columns per CTA -> 2 CTAs per SM, so
// one CTA's softmax overlaps the other's MMAs/TMA.
//
// Backward = attn_delta (delta = rowsum(dO o O)) + ONE kernel that recomputes S/P from the saved log-sum-exp (no
// T x T tensor ever reaches HBM): CTA = (key block, head, utterance) loops over query blocks,
//   dV += P^T dO, dK += dS^T Q accumulate in TMEM, dQ partials = dS K go to a per-key-block fp32 slab that
//   attn_dq_cast sums, with dS = P o (dP - delta) * scale and dP = dO V^T.  See attn_bwd_kernel below.
// The relative-position variant (REL) adds the shifted (q+v) Pw^T bias: fused in the forward, read back from a
// materialised bias in the backward.
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int AT_BQ = 128;   // query rows per tile
constexpr int AT_BK = 128;   // keys per block
constexpr int AT_DH = 64;
constexpr float LOG2E = 1.4426950408889634f;

// byte offset of element (row, col) inside a K-major SWIZZLE_128B bf16 tile made of 64-column blocks of
// [rows][128 B]; `rows` = rows per block.  16-byte chunk index is XORed with (row & 7).
__device__ __forceinline__ uint32_t sw128_offset(int row, int col, int rows) {
  const int blk = col >> 6, cc = col & 63;
  return static_cast<uint32_t>(blk * rows * 128 + row * 128 + ((((cc >> 3) ^ (row & 7)) << 4) | ((cc & 7) << 1)));
}

struct AttnParams {
  const int32_t* lengths;
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int B, T, H;
  int q_col, k_col, v_col;     // column offsets (elements) of head 0 inside a qkv row
  float scale_log2;            // scale * log2(e)
  int Tpad;                    // even row pitch of the dropout index space
  DropCfg drop;
  int qv_col;                  // REL: column of head 0 of (q + pos_bias_v) inside a qkv row
};

// Per-lane barrel shifter for the rel_shift: out[t] = in[t + s], s in [0, 32), all register resident (190 selects
// per 32 outputs).  tcgen05.ld hands every lane of a warp the SAME columns; the shift needs lane-dependent ones.
__device__ __forceinline__ void skew_select(const uint32_t (&in)[64], int s, uint32_t (&out)[32]) {
  uint32_t a[48], b[40], c[36], d[34];
  const bool s16 = s & 16, s8 = s & 8, s4 = s & 4, s2 = s & 2, s1 = s & 1;
#pragma unroll
  for (int t = 0; t < 48; ++t) a[t] = s16 ? in[t + 16] : in[t];
#pragma unroll
  for (int t = 0; t < 40; ++t) b[t] = s8 ? a[t + 8] : a[t];
#pragma unroll
  for (int t = 0; t < 36; ++t) c[t] = s4 ? b[t + 4] : b[t];
#pragma unroll
  for (int t = 0; t < 34; ++t) d[t] = s2 ? c[t + 2] : c[t];
#pragma unroll
  for (int t = 0; t < 32; ++t) out[t] = s1 ? d[t + 1] : d[t];
}

// 256 threads: thread (row = tid & 127, half = tid >> 7) owns query row `row` and 64 of the 128 keys of every key block
// (and 32 of the 64 output columns).  K is double-buffered and V is fetched one step ahead by TMA, so the only exposed
// waits are the two MMAs of the current block; 96 KB smem + 256 TMEM columns -> 2 CTAs per SM.
//
// REL = true adds the Transformer-XL relative-position term of HF Wav2Vec2ConformerSelfAttention._apply_relative_embeddings
// (modeling_wav2vec2_conformer.py:528-565; the reference only imports it):
//   score[i,j] = (q_i + u) . k_j + shift(BD)[i,j],   BD = (q + v) Pw^T,   shift(BD)[i,j] = BD[i, T-1-i+j]
// Q (column 0) already holds q + u and column qv_col holds q + v (the in-projection is run with augmented weights).
// For a (128-query, 128-key) tile only the 256 relative positions r in [r_lo, r_lo + 256), r_lo = T-128-q0+k0, occur:
// ONE extra UMMA  BD_win = Qv Pw[r_lo..]^T  (M128 N256) lands in TMEM cols [256,512) and row ii reads its scores from
// window columns 127-ii+jj -- the rel_shift is pure index math: a warp-uniform column base plus a per-lane barrel
// shift.  The combined score is written back over S (tcgen05.st) so the softmax below is shared.  512 TMEM columns
// -> one CTA per SM in this mode.
template <bool REL>
__global__ void __launch_bounds__(256, REL ? 1 : 2) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                   const __grid_constant__ CUtensorMap tmPW,
                                                                   const AttnParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  constexpr uint32_t OFF_Q = 0, OFF_K = 16384 /*2 bufs*/, OFF_V = 49152, OFF_P = 65536, OFF_X = 98304, OFF_BAR = 99328;
  constexpr uint32_t OFF_QV = 100352, OFF_PW = 116736 /*[256][64] window*/;     // REL only
  constexpr uint32_t TM_COLS = REL ? 512 : 256, TM_BD = 256;
  float* sx = reinterpret_cast<float*>(smem + OFF_X);          // [2][128] row-max / row-sum exchange between halves
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_k = bar_q + 1;                                 // [2]
  uint64_t* bar_v = bar_q + 3;
  uint64_t* bar_s = bar_q + 4;
  uint64_t* bar_o = bar_q + 5;
  uint64_t* bar_pw = bar_q + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 7);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.T;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1); mbar_init(&bar_k[0], 1); mbar_init(&bar_k[1], 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_pw, 1);
    fence_barrier_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, TM_COLS); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  pdl_wait();                    // barrier init + TMEM allocation above overlapped the previous kernel's tail
  const int len = min(p.lengths[b], p.T);
  const int nkb = (len + AT_BK - 1) / AT_BK;

  // REL: relative positions of key block j start at r_lo = T-128-q0 + 128 j (negative / past-the-end rows: TMA zero fill)
  auto load_pw = [&](int j) {
    const int r_lo = p.T - AT_BQ - q0 + j * AT_BK;
    mbar_expect_tx(bar_pw, 32768);
    tma_load_2d(smem + OFF_PW, &tmPW, bar_pw, h * AT_DH, r_lo);
    tma_load_2d(smem + OFF_PW + 16384, &tmPW, bar_pw, h * AT_DH, r_lo + 128);
  };
  if (tid == 0 && nkb > 0) {
    if constexpr (REL) tma_prefetch_desc(&tmPW);
    mbar_expect_tx(bar_q, REL ? 32768 : 16384);
    tma_load_2d(smem + OFF_Q, &tmQKV, bar_q, p.q_col + h * AT_DH, row_base + q0);
    if constexpr (REL) {
      tma_load_2d(smem + OFF_QV, &tmQKV, bar_q, p.qv_col + h * AT_DH, row_base + q0);
      load_pw(0);
    }
    mbar_expect_tx(&bar_k[0], 16384);
    tma_load_2d(smem + OFF_K, &tmQKV, &bar_k[0], p.k_col + h * AT_DH, row_base);
    mbar_expect_tx(bar_v, 16384);
    tma_load_2d(smem + OFF_V, &tmQKV, bar_v, p.v_col + h * AT_DH, row_base);
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
  constexpr uint32_t idesc_bd = umma_idesc_bf16(128, 256, false, false);
  // REL: first BD-window column this warp's lanes read for key 0 of the thread's half; lane l adds 31-l
  const uint32_t bd_col0 = TM_BD + 96 - 32 * (warp & 3) + 64 * half;

  float o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;                 // l_run: this thread's half of the row sum
  const DropCfg drop = drop_resolve(p.drop);
  const float inv_keep8 = 256.f / (256.f - static_cast<float>(drop.thresh16 >> 8));
  const int q_idx = q0 + row;                           // query position inside the utterance
  const uint64_t drop_row = (static_cast<uint64_t>(b * p.H + h) * p.T + q_idx) * p.Tpad;   // Tpad % 4 == 0

  for (int j = 0; j < nkb; ++j) {
    const uint32_t ph = j & 1, kb = j & 1;
    if (tid == 0) {
      if (j + 1 < nkb) {                                // K_{j+1} into the other buffer (its last reader, S_{j-1}, is done)
        mbar_expect_tx(&bar_k[kb ^ 1], 16384);
        tma_load_2d(smem + OFF_K + (kb ^ 1) * 16384, &tmQKV, &bar_k[kb ^ 1], p.k_col + h * AT_DH, row_base + (j + 1) * AT_BK);
      }
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[kb], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < AT_DH / 16; ++k)
        umma_bf16(tmem_base, umma_desc_sw128(base + OFF_Q + k * 32, 16, 1024),
                  umma_desc_sw128(base + OFF_K + kb * 16384 + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      if constexpr (REL) {
        mbar_wait(bar_pw, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_DH / 16; ++k)
          umma_bf16(tmem_base + TM_BD, umma_desc_sw128(base + OFF_QV + k * 32, 16, 1024),
                    umma_desc_sw128(base + OFF_PW + k * 32, 16, 1024), idesc_bd, k > 0 ? 1u : 0u);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if constexpr (REL) {
      if (tid == 0 && j + 1 < nkb) load_pw(j + 1);      // the product that read this window has retired
    }

    // ---- online softmax: this thread's 64 keys of its row ----
    // chunks that lie entirely below `len` (all but the utterance's last one) take a predicate-free path: the softmax
    // is issue-bound, every instruction per score counts
    const int key0 = j * AT_BK + half * 64;
    float m_loc = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tm_lane + half * 64 + c, v);
      if constexpr (REL) {
        uint32_t w[64], sh[32];
        tmem_ld32(tm_lane + bd_col0 + c, w);
        tmem_ld32(tm_lane + bd_col0 + c + 32, w + 32);
        tmem_wait_ld();
        skew_select(w, 31 - (tid & 31), sh);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(sh[i]));
        tmem_st32(tm_lane + half * 64 + c, v);          // combined score back over S: the exp pass below re-reads it
      } else {
        tmem_wait_ld();
      }
      if (key0 + c + 32 <= len) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m_loc = fmaxf(m_loc, __uint_as_float(v[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + c + i < len) m_loc = fmaxf(m_loc, __uint_as_float(v[i]));
      }
    }
    if constexpr (REL) tmem_wait_st();
    sx[half * 128 + row] = m_loc;
    __syncthreads();
    const float m_blk = fmaxf(m_loc, sx[(half ^ 1) * 128 + row]);
    const float m_new = fmaxf(m_run, m_blk * p.scale_log2);
    const float alpha = ex2_approx(m_run - m_new);      // first block: 2^(-inf) = 0
    float l_blk = 0.f;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tm_lane + half * 64 + c, v);
      tmem_wait_ld();
      float pr[32];
      if (key0 + c + 32 <= len) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
          l_blk += pr[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pr[i] = (key0 + c + i < len) ? ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new)) : 0.f;
          l_blk += pr[i];
        }
      }
      if (drop.thresh16 == 32768u) {
        const uint32_t bits = drop_bits32(drop, static_cast<uint64_t>(b * p.H + h) * p.T + q_idx, p.T, key0 + c);
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] *= (bits >> i) & 1u ? 2.f : 0.f;
      } else if (drop.thresh16) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float sc[4];
          drop_scale4(drop, drop_row + key0 + c + i, inv_keep8, sc);
          pr[i] *= sc[0]; pr[i + 1] *= sc[1]; pr[i + 2] *= sc[2]; pr[i + 3] *= sc[3];
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 u = make_uint4(pack_bf16(pr[g * 8], pr[g * 8 + 1]), pack_bf16(pr[g * 8 + 2], pr[g * 8 + 3]),
                                   pack_bf16(pr[g * 8 + 4], pr[g * 8 + 5]), pack_bf16(pr[g * 8 + 6], pr[g * 8 + 7]));
        *reinterpret_cast<uint4*>(smem + OFF_P + sw128_offset(row, half * 64 + c + g * 8, 128)) = u;
      }
    }
    l_run = l_run * alpha + l_blk;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] *= alpha;

    fence_proxy_async();          // P (generic-proxy stores) -> visible to the UMMA async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < AT_BK / 16; ++k)
        umma_bf16(tmem_base + 128, umma_desc_sw128(base + OFF_P + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  umma_desc_sw128(base + OFF_V + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < nkb) {                      // V_{j+1}: the PV product that read V_j has retired
      mbar_expect_tx(bar_v, 16384);
      tma_load_2d(smem + OFF_V, &tmQKV, bar_v, p.v_col + h * AT_DH, row_base + (j + 1) * AT_BK);
    }
    {
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tm_lane + 128 + half * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();              // sP and both TMEM regions are free for the next block
  }

  // combine the two halves' row sums.  len == 0 -> l == 0 -> 0/0 = NaN, what torch's softmax gives for an all-masked row
  sx[half * 128 + row] = l_run;
  __syncthreads();
  const float l_tot = l_run + sx[(half ^ 1) * 128 + row];
  if (q_idx < p.T) {
    const float inv_l = 1.f / l_tot;
    __nv_bfloat16* dst = p.out + static_cast<int64_t>(row_base + q_idx) * p.ld_out + h * AT_DH + half * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 u = make_uint4(pack_bf16(o[g * 8] * inv_l, o[g * 8 + 1] * inv_l),
                                 pack_bf16(o[g * 8 + 2] * inv_l, o[g * 8 + 3] * inv_l),
                                 pack_bf16(o[g * 8 + 4] * inv_l, o[g * 8 + 5] * inv_l),
                                 pack_bf16(o[g * 8 + 6] * inv_l, o[g * 8 + 7] * inv_l));
      *reinterpret_cast<uint4*>(dst + g * 8) = u;
    }
    if (p.lse && half == 0) p.lse[static_cast<int64_t>(b * p.H + h) * p.T + q_idx] = m_run + log2f(l_tot);
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

constexpr uint32_t ATTN_FWD_SMEM = 99328 + 64 + 1024;
constexpr uint32_t ATTN_FWD_REL_SMEM = 116736 + 32768 + 1024;
// ================================================================================================
// backward
// ================================================================================================
struct AttnBwdParams {
  const int32_t* lengths;
  const float* lse;        // [B,H,T] log2-domain log-sum-exp saved by the forward pass
  const float* delta;      // [B,H,T] rowsum(dO o O)
  __nv_bfloat16* dqkv;     // [B*T, ld] laid out like qkv
  int64_t ld;
  int B, T, H;
  int q_col, k_col, v_col;
  float scale, scale_log2;
  int Tpad;
  DropCfg drop;
  void* delta_dbg;         // only used by -DB2C_ATTN_TIMING builds (cycle stamps of one CTA)
  // REL (relative-position scores): bd = (q+v) Pw^T fp32 [B*T, H, Rp] (un-shifted, column r = relative row index);
  // dbd receives dS at the shifted position (bf16, same layout, zeroed by the caller) -- see b2c_attention_rel_bwd
  const float* bd;
  __nv_bfloat16* dbd;
  int64_t Rp;
};

// delta[b,h,t] = sum_c dO[row, h*64+c] * O[row, h*64+c].  Eight lanes share one (row, head): each reads ONE 16-byte piece of
// O and of dO (a warp instruction covers 4 x 128 contiguous bytes), three shuffles finish the dot product.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                        const __nv_bfloat16* __restrict__ dout, int64_t ld_out,
                                                        float* __restrict__ delta, int B, int T, int H) {
  pdl_trigger();
  pdl_wait();
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int part = static_cast<int>(g & 7);
  const int64_t idx = g >> 3;                                                          // (row, h)
  const bool live = idx < static_cast<int64_t>(B) * T * H;
  const int h = static_cast<int>(idx % H);
  const int64_t row = live ? idx / H : 0;
  float acc = 0.f;
  if (live) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * ld_out + h * AT_DH + part * 8);
    const uint4 d = *reinterpret_cast<const uint4*>(dout + row * ld_out + h * AT_DH + part * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16(aw[j]), y = unpack_bf16(dw[j]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && part == 0) {
    const int b_ = static_cast<int>(row / T), t = static_cast<int>(row % T);
    delta[(static_cast<int64_t>(b_) * H + h) * T + t] = acc;
  }
}

// Thread (row = tid & 127, column quarter cq = tid >> 7) turns 32 columns of its S row and dP row (TMEM) into
// P (optionally dropped, feeds dV) and dS = P o (dP*drop - delta) * scale, both written as bf16 [q][key] K-major
// SW128 tiles (2 x [128][64]).  Rows past the utterance end contribute nothing: with lse = +inf their probabilities
// are exactly 2^(-inf) = 0, no per-element predicate needed (lse_r is set by the caller).
// REL: the relative-position score term is read from the materialised BD at the rel_shift index r = T-1-q+key and dS
// is stored to dBD at the same index (the two products that consume dBD are plain GEMMs, see b2c_attention_rel_bwd).
template <bool REL>
__device__ __forceinline__ void attn_bwd_tile_math(const AttnBwdParams& p, uint32_t tm_s, uint32_t tm_dp, uint8_t* sP,
                                                   uint8_t* sdS, int row, int cq, int key0, int len, float lse_r,
                                                   float delta_r, uint64_t drop_row, int64_t rel_row, int q_idx,
                                                   uint64_t* bar_sfree, uint64_t row_lin) {
  const DropCfg drop = drop_resolve(p.drop);
  const float inv_keep8 = 256.f / (256.f - static_cast<float>(drop.thresh16 >> 8));
  const int c = cq * 32;
  uint32_t vs[32], vd[32];
  __syncwarp();
  tmem_ld32(tm_s + c, vs);
  tmem_ld32(tm_dp + c, vd);
  tmem_wait_ld();
  // S / dP now live in registers: the issuer may overwrite the accumulators with the next block's products while this
  // block's ~4 k cycles of tile math run
  tc_fence_before();
  mbar_arrive(bar_sfree);
  const bool full = key0 + c + 32 <= len;
  const bool half_p = drop.thresh16 == 32768u;
  const uint32_t bits = half_p ? drop_bits32(drop, row_lin, p.T, key0 + c) : 0u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pr[8], ds[8];
    float bb[8];
    // rel_row >= 0 only for real query rows; keys past T have no relative position (and no probability mass)
    const int64_t rel_at = rel_row + (p.T - 1 - q_idx) + key0 + c + g * 8;
    if constexpr (REL) {
#pragma unroll
      for (int e = 0; e < 8; ++e) bb[e] = (rel_row >= 0 && key0 + c + g * 8 + e < p.T) ? __ldg(p.bd + rel_at + e) : 0.f;
    }
#pragma unroll
    for (int i4 = 0; i4 < 8; i4 += 4) {
      const int i = g * 8 + i4;
      float dsc[4] = {1.f, 1.f, 1.f, 1.f};
      if (half_p) {
#pragma unroll
        for (int e = 0; e < 4; ++e) dsc[e] = (bits >> (i + e)) & 1u ? 2.f : 0.f;
      } else if (drop.thresh16) {
        drop_scale4(drop, drop_row + key0 + c + i, inv_keep8, dsc);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (REL) vs[i + e] = __float_as_uint(__uint_as_float(vs[i + e]) + bb[i4 + e]);
        float pv = ex2_approx(fmaf(__uint_as_float(vs[i + e]), p.scale_log2, -lse_r));
        if (!full && key0 + c + i + e >= len) pv = 0.f;
        pr[i4 + e] = pv * dsc[e];
        ds[i4 + e] = (pv * p.scale) * fmaf(__uint_as_float(vd[i + e]), dsc[e], -delta_r);
      }
    }
    const uint32_t off = sw128_offset(row, c + g * 8, 128);
    *reinterpret_cast<uint4*>(sP + off) = make_uint4(pack_bf16(pr[0], pr[1]), pack_bf16(pr[2], pr[3]),
                                                     pack_bf16(pr[4], pr[5]), pack_bf16(pr[6], pr[7]));
    *reinterpret_cast<uint4*>(sdS + off) = make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]),
                                                      pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
    if constexpr (REL) {
      if (rel_row >= 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (key0 + c + g * 8 + e < p.T) p.dbd[rel_at + e] = __float2bfloat16_rn(ds[e]);
      }
    }
  }
}

// One kernel for dQ, dK, dV.  CTA = (128-key block j, head, utterance), 16 math warps + 1 issuer warp, loops over the
// utterance's query blocks i with Q_i / dO_i double-buffered by TMA:
//   UMMA  S = Q_i K_j^T, dP = dO_i V_j^T                  -> TMEM [0,128) and [128,256)
//   regs  P = exp2(S*scale - lse), dS = P o (dP - delta)   -> smem (bf16, K-major SW128, DOUBLE-buffered), dropout regenerated
//   UMMA  dV += P^T dO_i, dK += dS^T Q_i  (accumulate over i in TMEM), dQp = dS K_j (fresh)
//   regs  dQp -> smem -> TMA REDUCE-ADD into the one fp32 dQ buffer; attn_dq_cast_kernel converts it -> bf16
// S and dP are recomputed here from the saved log-sum-exp, so no T x T tensor ever reaches HBM.
//
// Warp specialisation: the 512 math threads (four per tile row, 32 columns each: the tile math is latency bound) never
// issue an MMA.  Warp 16's elected lane owns TMA and tcgen05.mma: when the math threads have published P/dS of block i
// (mbarrier, 512 arrivals) it issues S/dP of block i+1 FIRST and then the three products of block i, so the math warps
// are back at work on block i+1 after ~1 k cycles while block i's 24 UMMAs (2.5 k cycles to issue from one thread +
// 1 k to run -- previously spent with all 512 threads parked on a barrier) drain underneath; dQp of block i is read out
// after the tile math of block i+1.  P/dS are double-buffered so block i+1's tile math does not wait for them.
template <bool REL>
__global__ void __launch_bounds__(544, 1) attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                          const __grid_constant__ CUtensorMap tmDO,
                                                          const __grid_constant__ CUtensorMap tmDQ,
                                                          const __grid_constant__ CUtensorMap tmDKV,
                                                          const AttnBwdParams p, float* __restrict__ dq_f32) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  constexpr uint32_t OFF_K = 0, OFF_V = 16384, OFF_Q = 32768 /*2 bufs*/, OFF_DO = 65536 /*2 bufs*/, OFF_P = 98304 /*2 bufs*/,
                     OFF_DS = 163840 /*2 bufs*/, OFF_BAR = 229376;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_q = bar_kv + 1;            // [2] Q_i / dO_i landed
  uint64_t* bar_s = bar_kv + 3;            // S / dP of the next block are in TMEM
  uint64_t* bar_acc = bar_kv + 4;          // dV / dK / dQp products of a block have retired
  uint64_t* bar_tile = bar_kv + 5;         // [2] P / dS buffer published by the 512 math threads
  uint64_t* bar_sfree = bar_kv + 7;        // S / dP accumulators read out by the 512 math threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool issuer_warp = warp == 16;
#ifdef B2C_ATTN_TIMING
  // stamps live BEHIND the real partial-dQ slabs: a timing run allocates one extra slab (tools/attn_timing.py)
  long long* dbg = reinterpret_cast<long long*>(static_cast<float*>(p.delta_dbg) +
                                                static_cast<size_t>((p.T + AT_BK - 1) / AT_BK) * p.B * p.T * p.H * AT_DH);
#define B2C_TK(slot) if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 40) dbg[slot] = clock64();
#define B2C_TM(i_, slot) if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 40) dbg[16 + (i_) * 8 + (slot)] = clock64();
#define B2C_TI(i_, slot) if (blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 40) dbg[64 + (i_) * 8 + (slot)] = clock64();
#else
#define B2C_TK(slot)
#define B2C_TM(i_, slot)
#define B2C_TI(i_, slot)
#endif
  B2C_TK(0)
  const int row = tid & 127, cq = (tid >> 7) & 3;    // cq: which 32 tile columns / 16 head-dim columns this thread owns
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = min(p.lengths[b], p.T);
  const int key0 = j * AT_BK, row_base = b * p.T;
  const int key_idx = key0 + row;
  // this thread stores 16 of the 64 head-dim columns of its key row
  __nv_bfloat16* dk_dst = key_idx < p.T ? p.dqkv + static_cast<int64_t>(row_base + key_idx) * p.ld + p.k_col + h * AT_DH + cq * 16 : nullptr;
  __nv_bfloat16* dv_dst = key_idx < p.T ? p.dqkv + static_cast<int64_t>(row_base + key_idx) * p.ld + p.v_col + h * AT_DH + cq * 16 : nullptr;
  if (key0 >= len) {                       // every key of this block is masked: P = 0 -> dK = dV = 0, dQ partial = 0
    if (issuer_warp) return;
    // (no dQ contribution: the dQ buffer is zero-initialised and every live key block ADDS its partial tile to it)
    if (dk_dst) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        reinterpret_cast<uint4*>(dk_dst)[g] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(dv_dst)[g] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ); tma_prefetch_desc(&tmDKV);
    mbar_init(bar_kv, 1); mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1); mbar_init(bar_s, 1); mbar_init(bar_acc, 1);
    mbar_init(&bar_tile[0], 512); mbar_init(&bar_tile[1], 512); mbar_init(bar_sfree, 512);
    fence_barrier_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  B2C_TK(1)
  const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 320, TM_DQ = 384;
  const int nqb = (p.T + AT_BQ - 1) / AT_BQ;

  if (issuer_warp) {
    // ================================ TMA + MMA issuer (one elected lane) ================================
    if ((tid & 31) == 0) {
      auto load_q = [&](int i) {
        const int bf = i & 1;
        mbar_expect_tx(&bar_q[bf], 32768);
        tma_load_2d(smem + OFF_Q + bf * 16384, &tmQKV, &bar_q[bf], p.q_col + h * AT_DH, row_base + i * AT_BQ);
        tma_load_2d(smem + OFF_DO + bf * 16384, &tmDO, &bar_q[bf], h * AT_DH, row_base + i * AT_BQ);
      };
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_mm = umma_idesc_bf16(128, 64, true, true);
      constexpr uint32_t idesc_km = umma_idesc_bf16(128, 64, false, true);
      auto issue_s_dp = [&](int i) {       // S = Q_i K^T and dP = dO_i V^T
        const uint32_t bf = i & 1;
        const uint32_t q_addr = base + OFF_Q + bf * 16384, do_addr = base + OFF_DO + bf * 16384;
        mbar_wait(&bar_q[bf], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + TM_S, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                    umma_desc_sw128(base + OFF_K + k * 32, 16, 1024), idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + TM_DP, umma_desc_sw128(do_addr + k * 32, 16, 1024),
                    umma_desc_sw128(base + OFF_V + k * 32, 16, 1024), idesc_kk, k > 0);
        tc_commit(bar_s);
      };
      mbar_expect_tx(bar_kv, 32768);
      tma_load_2d(smem + OFF_K, &tmQKV, bar_kv, p.k_col + h * AT_DH, row_base + key0);
      tma_load_2d(smem + OFF_V, &tmQKV, bar_kv, p.v_col + h * AT_DH, row_base + key0);
      load_q(0);
      if (nqb > 1) load_q(1);
      mbar_wait(bar_kv, 0);
      issue_s_dp(0);
      for (int i = 0; i < nqb; ++i) {
        const uint32_t bf = i & 1;
        const uint32_t q_addr = base + OFF_Q + bf * 16384, do_addr = base + OFF_DO + bf * 16384;
        const uint32_t p_addr = base + OFF_P + bf * 32768, ds_addr = base + OFF_DS + bf * 32768;
        B2C_TI(i, 0)
        if (i + 1 < nqb) {                   // S/dP of block i are in registers: the next block's products go first
          mbar_wait(bar_sfree, i & 1);
          tc_fence_after();
          issue_s_dp(i + 1);
        }
        B2C_TI(i, 1)
        // P/dS of block i published; dQp of block i-1 drained
        mbar_wait(&bar_tile[bf], (i >> 1) & 1);
        tc_fence_after();
        B2C_TI(i, 2)
#pragma unroll
        for (int k = 0; k < 8; ++k)          // dV += P^T dO   (reduction over the 128 queries)
          umma_bf16(tmem_base + TM_DV, umma_desc_sw128(p_addr + k * 2048, 16384, 1024),
                    umma_desc_sw128(do_addr + k * 2048, 8192, 1024), idesc_mm, (i | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)          // dK += dS^T Q
          umma_bf16(tmem_base + TM_DK, umma_desc_sw128(ds_addr + k * 2048, 16384, 1024),
                    umma_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_mm, (i | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)          // dQp = dS K     (reduction over the 128 keys; K tile read MN-major)
          umma_bf16(tmem_base + TM_DQ, umma_desc_sw128(ds_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(base + OFF_K + k * 2048, 8192, 1024), idesc_km, k > 0);
        tc_commit(bar_acc);
        B2C_TI(i, 3)
        if (i + 2 < nqb) {                   // this block's Q/dO buffer is free once its products have retired
          mbar_wait(bar_acc, i & 1);
          load_q(i + 2);
        }
        B2C_TI(i, 4)
      }
    }
  } else {
    // ================================ 512 math threads ================================
    // dQp of block i: TMEM -> registers -> the P buffer of block i (its readers, the dV products, have retired) as two
    // 128B-swizzled [128 rows][32 fp32] boxes -> ONE thread issues two TMA stores into this key block's partial slab.
    // With ~230 KB of shared memory there is no L1 left and the per-thread 64-byte row pieces cost 1.8 k cycles per
    // block as plain stores; the TMA unit also clips the rows past the utterance end (3-D map: column, frame, slab*B+b).
    auto drain_dq = [&](int i) {
      mbar_wait(bar_acc, i & 1);
      tc_fence_after();
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tm_lane + TM_DQ + cq * 16, v);
      tmem_wait_ld();
      uint8_t* stg = smem + OFF_P + (i & 1) * 32768 + (cq >> 1) * 16384 + row * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(stg + ((((cq & 1) * 4 + g) ^ (row & 7)) << 4)) =
            make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      fence_proxy_async();
      named_barrier_sync(1, 512);            // all 512 math threads have staged their pieces
      if (tid == 0) {
        const uint8_t* src = smem + OFF_P + (i & 1) * 32768;
        tma_reduce_add_3d(&tmDQ, src, h * AT_DH, i * AT_BQ, b);            // dQ[b, i-th query block, head h] += dS K_j
        tma_reduce_add_3d(&tmDQ, src + 16384, h * AT_DH + 32, i * AT_BQ, b);
        tma_store_commit();
      }
    };
    for (int i = 0; i < nqb; ++i) {
      const uint32_t ph = i & 1, bf = i & 1;
      const int q_idx = i * AT_BQ + row;
      const int64_t stat = static_cast<int64_t>(b * p.H + h) * p.T + min(q_idx, p.T - 1);
      const float lse_r = q_idx < p.T ? p.lse[stat] : INFINITY, delta_r = q_idx < p.T ? p.delta[stat] : 0.f;
      // (the products of block i-2 that read this P/dS buffer have retired: drain_dq(i-2) waited for them)
      B2C_TM(i, 0)
      if (i >= 2) {                          // P[i & 1] was the staging buffer of dQp(i-2): its TMA stores must have read it
        if (tid == 0) tma_store_wait_read();
        named_barrier_sync(1, 512);
      }
      mbar_wait(bar_s, ph);
      tc_fence_after();
      B2C_TM(i, 1)
      attn_bwd_tile_math<REL>(p, tm_lane + TM_S, tm_lane + TM_DP, smem + OFF_P + bf * 32768, smem + OFF_DS + bf * 32768, row,
                              cq, key0, len, lse_r, delta_r, (static_cast<uint64_t>(b * p.H + h) * p.T + q_idx) * p.Tpad,
                              (REL && q_idx < p.T) ? (static_cast<int64_t>(row_base + q_idx) * p.H + h) * p.Rp : -1, q_idx,
                              bar_sfree, static_cast<uint64_t>(b * p.H + h) * p.T + q_idx);
      B2C_TM(i, 2)
      if (i > 0) drain_dq(i - 1);            // block i-1's products ran underneath this block's tile math
      B2C_TM(i, 3)
      fence_proxy_async();                   // P/dS (generic-proxy stores) -> visible to the UMMA async proxy
      tc_fence_before();
      mbar_arrive(&bar_tile[bf]);
      B2C_TM(i, 4)
    }
    drain_dq(nqb - 1);
    if (tid == 0) tma_store_wait_read();     // shared memory stays valid until the last stores have read it
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  B2C_TK(3)
  if (!issuer_warp) {
    // dV / dK: TMEM -> bf16 -> the (now idle) V / K tiles as 128B-swizzled [128 keys][64] boxes -> two TMA stores that
    // clip at the utterance end (the per-thread 32-byte row pieces were as slow here as for dQ)
    uint32_t v[16];
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __syncwarp();
      tmem_ld16(tm_lane + (which ? TM_DK : TM_DV) + cq * 16, v);
      tmem_wait_ld();
      uint8_t* stg = smem + (which ? OFF_K : OFF_V) + row * 128;
#pragma unroll
      for (int g = 0; g < 2; ++g)
        *reinterpret_cast<uint4*>(stg + (((cq * 2 + g) ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16(__uint_as_float(v[g * 8]), __uint_as_float(v[g * 8 + 1])),
                       pack_bf16(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])),
                       pack_bf16(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])),
                       pack_bf16(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])));
    }
    fence_proxy_async();
    named_barrier_sync(1, 512);
    if (tid == 0) {
      tma_store_3d(&tmDKV, smem + OFF_V, p.v_col + h * AT_DH, key0, b);
      tma_store_3d(&tmDKV, smem + OFF_K, p.k_col + h * AT_DH, key0, b);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  B2C_TK(4)
  if (warp == 0) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
  B2C_TK(5)
}

// dQ = sum over key blocks of the fp32 partial slabs [nkb][rows, H*64] -> the Q slab of dqkv (bf16, pitch ld)
__global__ void __launch_bounds__(256) attn_dq_cast_kernel(const float* __restrict__ dq, __nv_bfloat16* __restrict__ dqkv,
                                                          int64_t rows, int cols, int64_t ld, int q_col, int nkb) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t n = rows * cols;
  if (i >= n) return;
  const int64_t r = i / cols;
  const int c = static_cast<int>(i % cols);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
  for (int k = 0; k < nkb; ++k) {
    const float4 x = *reinterpret_cast<const float4*>(dq + k * n + i), y = *reinterpret_cast<const float4*>(dq + k * n + i + 4);
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    b4.x += y.x; b4.y += y.y; b4.z += y.z; b4.w += y.w;
  }
  *reinterpret_cast<uint4*>(dqkv + r * ld + q_col + c) =
      make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b4.x, b4.y), pack_bf16(b4.z, b4.w));
}

constexpr uint32_t ATTN_BWD_SMEM = 229376 + 128 + 1024;

}  // namespace b2c

using namespace b2c;

static int attention_fwd_launch(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, int64_t qv_off, const void* pw,
                                int64_t ld_pw, int64_t pw_rows, const int32_t* lengths, void* out, int64_t ld_out,
                                float* lse, int B, int T, int H, float scale, float drop_p, uint64_t drop_seed,
                                void* stream) {
  const bool rel = pw != nullptr;
  B2C_REQUIRE(qkv && lengths && out, "b2c_attention_fwd: null pointer");
  B2C_REQUIRE(B > 0 && T > 0 && H > 0, "b2c_attention_fwd: empty problem");
  B2C_REQUIRE(ld % 8 == 0 && ld_out % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && qv_off % 8 == 0,
              "b2c_attention_fwd: pitches/offsets must be multiples of 8 elements");
  B2C_REQUIRE(static_cast<int64_t>(B) * T < (1ll << 31), "b2c_attention_fwd: B*T too large");
  CUtensorMap tm, tp;
  int rc = get_tensor_map_2d(&tm, qkv, 2, static_cast<uint64_t>(ld), static_cast<uint64_t>(B) * T, ld * 2, 64, 128, true);
  if (rc) return rc;
  if (rel) {
    B2C_REQUIRE(ld_pw % 8 == 0 && ld_pw >= static_cast<int64_t>(H) * AT_DH && pw_rows >= 2ll * T - 1,
                "b2c_attention_rel_fwd: pw must be bf16 [>= 2T-1, ld_pw >= H*64], ld_pw a multiple of 8");
    rc = get_tensor_map_2d(&tp, pw, 2, static_cast<uint64_t>(H) * AT_DH, static_cast<uint64_t>(pw_rows), ld_pw * 2, 64, 128, true);
    if (rc) return rc;
  } else {
    tp = tm;
  }
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_FWD_SMEM));
    B2C_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_FWD_REL_SMEM));
    configured = true;
  }
  AttnParams p;
  p.lengths = lengths; p.out = static_cast<__nv_bfloat16*>(out); p.lse = lse; p.ld_out = ld_out;
  p.B = B; p.T = T; p.H = H;
  p.q_col = 0; p.k_col = static_cast<int>(k_off); p.v_col = static_cast<int>(v_off); p.qv_col = static_cast<int>(qv_off);
  p.scale_log2 = scale * LOG2E;
  p.Tpad = (T + 3) & ~3;
  p.drop = make_drop(drop_p, drop_seed);
  dim3 grid((T + AT_BQ - 1) / AT_BQ, H, B);
  if (rel) B2C_CUDA(launch_k(attn_fwd_kernel<true>, dim3(grid), dim3(256), ATTN_FWD_REL_SMEM, static_cast<cudaStream_t>(stream), tm, tp, p));
  else B2C_CUDA(launch_k(attn_fwd_kernel<false>, dim3(grid), dim3(256), ATTN_FWD_SMEM, static_cast<cudaStream_t>(stream), tm, tp, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_attention_fwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                                 void* out, int64_t ld_out, float* lse, int B, int T, int H, float scale, float drop_p,
                                 uint64_t drop_seed, void* stream) {
  return attention_fwd_launch(qkv, ld, k_off, v_off, 0, nullptr, 0, 0, lengths, out, ld_out, lse, B, T, H, scale, drop_p,
                              drop_seed, stream);
}

extern "C" int b2c_attention_rel_fwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, int64_t qv_off,
                                     const void* pw, int64_t ld_pw, int64_t pw_rows, const int32_t* lengths, void* out,
                                     int64_t ld_out, float* lse, int B, int T, int H, float scale, float drop_p,
                                     uint64_t drop_seed, void* stream) {
  B2C_REQUIRE(pw, "b2c_attention_rel_fwd: null pw");
  return attention_fwd_launch(qkv, ld, k_off, v_off, qv_off, pw, ld_pw, pw_rows, lengths, out, ld_out, lse, B, T, H, scale,
                              drop_p, drop_seed, stream);
}

static int attention_bwd_launch(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                                const void* out, const void* dout, int64_t ld_out, const float* lse, float* delta,
                                float* dq_f32, void* dqkv, const float* bd, void* dbd, int64_t Rp, int B, int T, int H,
                                float scale, float drop_p, uint64_t drop_seed, void* stream) {
  B2C_REQUIRE(qkv && lengths && out && dout && lse && delta && dq_f32 && dqkv, "b2c_attention_bwd: null pointer");
  B2C_REQUIRE(B > 0 && T > 0 && H > 0, "b2c_attention_bwd: empty problem");
  B2C_REQUIRE(ld % 8 == 0 && ld_out % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
              "b2c_attention_bwd: pitches/offsets must be multiples of 8 elements");
  const bool rel = bd != nullptr;
  B2C_REQUIRE(!rel || (dbd && Rp >= 2ll * T - 1), "b2c_attention_rel_bwd: dbd missing or Rp < 2T-1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tq, td;
  if (int rc = get_tensor_map_2d(&tq, qkv, 2, static_cast<uint64_t>(ld), static_cast<uint64_t>(B) * T, ld * 2, 64, 128, true)) return rc;
  if (int rc = get_tensor_map_2d(&td, dout, 2, static_cast<uint64_t>(H) * AT_DH, static_cast<uint64_t>(B) * T, ld_out * 2, 64, 128, true)) return rc;
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_BWD_SMEM));
    B2C_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_BWD_SMEM));
    configured = true;
  }
  const int64_t rows = static_cast<int64_t>(B) * T;
  const int64_t n = rows * H;
  B2C_CUDA(launch_k(attn_delta_kernel, dim3(static_cast<unsigned>((n * 8 + 255) / 256)), dim3(256), 0, st, 
      static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), ld_out, delta, B, T, H));
  AttnBwdParams p;
  p.lengths = lengths; p.lse = lse; p.delta = delta; p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld = ld;
  p.B = B; p.T = T; p.H = H;
  p.q_col = 0; p.k_col = static_cast<int>(k_off); p.v_col = static_cast<int>(v_off);
  p.scale = scale; p.scale_log2 = scale * LOG2E;
  p.Tpad = (T + 3) & ~3;
  p.drop = make_drop(drop_p, drop_seed);
  p.delta_dbg = dq_f32;     // timing builds scribble cycle stamps over the head of the dQ workspace
  p.bd = bd; p.dbd = static_cast<__nv_bfloat16*>(dbd); p.Rp = Rp;
  dim3 grid((T + AT_BK - 1) / AT_BK, H, B);
  // dQ accumulates in ONE zeroed fp32 buffer [b][frame][H*64]: every key-block CTA adds its partial tile with a TMA
  // reduce (round 1 wrote one slab per key block and summed them in a second kernel: 2/3 of that traffic is gone);
  // the frame dimension clips at T per utterance
  B2C_CUDA(cudaMemsetAsync(dq_f32, 0, static_cast<size_t>(B) * T * H * AT_DH * sizeof(float), st));
  CUtensorMap tdq;
  if (int rc = get_tensor_map_3d(&tdq, dq_f32, 4, static_cast<uint64_t>(H) * AT_DH, static_cast<uint64_t>(T),
                                 static_cast<uint64_t>(B), static_cast<uint64_t>(H) * AT_DH * 4,
                                 static_cast<uint64_t>(T) * H * AT_DH * 4, 32, 128, true))
    return rc;
  CUtensorMap tdkv;  // dqkv as [b][frame][ld] bf16 (dK / dV tiles are stored through it, clipped at T)
  if (int rc = get_tensor_map_3d(&tdkv, dqkv, 2, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(B),
                                 static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(T) * ld * 2, 64, 128, true))
    return rc;
  if (rel) B2C_CUDA(launch_k(attn_bwd_kernel<true>, dim3(grid), dim3(544), ATTN_BWD_SMEM, st, tq, td, tdq, tdkv, p, dq_f32));
  else B2C_CUDA(launch_k(attn_bwd_kernel<false>, dim3(grid), dim3(544), ATTN_BWD_SMEM, st, tq, td, tdq, tdkv, p, dq_f32));
  const int64_t n8 = rows * H * AT_DH / 8;
  B2C_CUDA(launch_k(attn_dq_cast_kernel, dim3(static_cast<unsigned>((n8 + 255) / 256)), dim3(256), 0, st, dq_f32, static_cast<__nv_bfloat16*>(dqkv),
                                                                               rows, H * AT_DH, ld, 0, 1));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_attention_bwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                                 const void* out, const void* dout, int64_t ld_out, const float* lse, float* delta,
                                 float* dq_f32, void* dqkv, int B, int T, int H, float scale, float drop_p,
                                 uint64_t drop_seed, void* stream) {
  return attention_bwd_launch(qkv, ld, k_off, v_off, lengths, out, dout, ld_out, lse, delta, dq_f32, dqkv, nullptr,
                              nullptr, 0, B, T, H, scale, drop_p, drop_seed, stream);
}

extern "C" int b2c_attention_rel_bwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                                     const void* out, const void* dout, int64_t ld_out, const float* lse, float* delta,
                                     float* dq_f32, void* dqkv, const float* bd, void* dbd, int64_t Rp, int B, int T,
                                     int H, float scale, float drop_p, uint64_t drop_seed, void* stream) {
  B2C_REQUIRE(bd && dbd, "b2c_attention_rel_bwd: null bd / dbd");
  return attention_bwd_launch(qkv, ld, k_off, v_off, lengths, out, dout, ld_out, lse, delta, dq_f32, dqkv, bd, dbd, Rp, B,
                              T, H, scale, drop_p, drop_seed, stream);
}
