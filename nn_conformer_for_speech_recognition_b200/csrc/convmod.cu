// Conformer convolution-module chain: GLU -> depthwise conv1d ('same', cross-correlation) -> BatchNorm1d -> SiLU
// (torchaudio/models/conformer.py:52-65), token-major, HBM/L2-bound.
//
// The reference runs 4 library kernels plus transposes on a (B,D,T) layout (GLU 3d, depthwise 2d, BN 2d, SiLU 2d
// elements of traffic per token = 9d).  Here the eval-mode chain is ONE kernel reading 2d and writing d bf16
// elements per token (algorithmic bytes 3*d*2 per token); train mode needs the batch statistics first, so it is
// two passes (5*d*2 bytes per token).
//
// Tiling: CTA = (utterance b, 128-frame time tile, 64-channel slab), 256 threads = 8 warps.  A lane owns two
// adjacent channels (bf16x2 / float2 accesses, 128..256 B coalesced per warp instruction); the gated activations for
// frames [t0-(k-1)/2, t0+128+(k-1)/2) are staged once in shared memory as fp32 [frames][64] (conflict-free: lanes
// read consecutive 8-byte words), each warp then produces 16 frames for its 64 channels as two groups of 8 outputs
// from a register sliding window (8+k-1 smem reads per 8 outputs instead of 8k), taps held in registers.
// Frames past `lengths[b]` are NOT zeroed (conformer.py:85-88); only the utterance edges are zero padded.
#include <stdlib.h>

#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int CM_TT = 128;     // output frames per CTA
constexpr int CM_CH = 64;      // channels per CTA
constexpr int CM_MAXK = 33;    // taps (repo default k=33, Conformer S/M/L k=31)
constexpr int CM_G = 8;        // outputs per register window (backward)
constexpr int CM_GF = 16;      // outputs per register window (forward: one window per warp)

constexpr int CM_UN = 10;      // row pairs per thread: 16 rows per CTA step x 10 = 160 >= 128 + (CM_MAXK - 1) - 1 rows
static_assert(16 * CM_UN >= CM_TT + CM_MAXK - 1, "one batch must cover the tile and its halo");
// Loads are UNCONDITIONAL from a clamped (always valid) address and the zero padding is applied when the value is
// consumed: a predicated load into a pre-zeroed register made the compiler put a MOV behind every LDG, and a MOV that
// reads a load's destination blocks in-order issue until the data is back -- the whole batch then ran one memory
// latency after the other (ncu: long-scoreboard stalls on MOV, 47 % of the samples).
__device__ __forceinline__ uint32_t glu_load(const __nv_bfloat16* __restrict__ u, uint2 (&av)[CM_UN], uint2 (&gv)[CM_UN],
                                             int b, int T, int d, int c0, int t_lo, int nrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & 15, rsel = lane >> 4;
  const int ch = c0 + hl * 4;
  const bool ch_ok = ch < d;                       // d % 8 == 0 and c0 % 64 == 0: a 4-channel group is all in or all out
  const __nv_bfloat16* base = u + static_cast<int64_t>(b) * T * (2 * d) + (ch_ok ? ch : c0);
  uint32_t okmask = 0;
#pragma unroll
  for (int q = 0; q < CM_UN; ++q) {
    const int r = warp * 2 + rsel + q * 16, t = t_lo + r;
    if (r < nrow && t >= 0 && t < T && ch_ok) okmask |= 1u << q;
    const __nv_bfloat16* row = base + static_cast<int64_t>(min(max(t, 0), T - 1)) * (2 * d);
    av[q] = *reinterpret_cast<const uint2*>(row);
    gv[q] = *reinterpret_cast<const uint2*>(row + d);
  }
  return okmask;
}
__device__ __forceinline__ void glu_store(const uint2 (&av)[CM_UN], const uint2 (&gv)[CM_UN], uint32_t okmask,
                                          float2* __restrict__ sG, int nrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = lane & 15, rsel = lane >> 4;
  float4* sG4 = reinterpret_cast<float4*>(sG);
#pragma unroll
  for (int q = 0; q < CM_UN; ++q) {
    const int r = warp * 2 + rsel + q * 16;
    if (r < nrow) {
      const float2 a0 = unpack_bf16(av[q].x), a1 = unpack_bf16(av[q].y);
      const float2 g0 = unpack_bf16(gv[q].x), g1 = unpack_bf16(gv[q].y);
      const float keep = (okmask >> q) & 1u ? 1.f : 0.f;                                 // zero padding / channel tail
      sG4[r * 16 + hl] = make_float4(keep * a0.x * sigmoid_f(g0.x), keep * a0.y * sigmoid_f(g0.y),
                                     keep * a1.x * sigmoid_f(g1.x), keep * a1.y * sigmoid_f(g1.y));
    }
  }
}
// stage GLU(u) for frames [t_lo, t_lo + nrow) of utterance b into smem (zero outside [0,T)) as fp32 [row][64].
// A half-warp takes one row (lane = 4 adjacent channels: one 8-byte load from each GLU half, one conflict-free
// 16-byte smem store), so a warp instruction covers two rows and the CTA 16; ALL loads of the tile are issued before
// the first is consumed (one memory round trip per CTA; the row-at-a-time version exposed one per row).
__device__ __forceinline__ void stage_glu(const __nv_bfloat16* __restrict__ u, float2* __restrict__ sG, int b, int T,
                                          int d, int c0, int t_lo, int nrow) {
  uint2 av[CM_UN], gv[CM_UN];
  const uint32_t okmask = glu_load(u, av, gv, b, T, d, c0, t_lo, nrow);
  glu_store(av, gv, okmask, sG, nrow);
}

// depthwise taps of this CTA's 64 channels -> smem, tap-major [k][64] (coalesced global read, conflict-free float2 reads)
constexpr int CM_WP = CM_CH + 2;   // tap-row pitch in shared memory: 66 floats (consecutive taps of a channel land 2 banks apart: the
                                   // transposing store is 2-way conflicted instead of 31-way; float2 reads stay 8-byte aligned)
constexpr int CM_RS = CM_MAXK + 3;   // reduction slots per channel in the backward kernel: k taps, dbdw, d(pw1 bias) of both GLU halves
constexpr int CM_TAPLD = (CM_CH * CM_MAXK + 255) / 256;    // tap elements per thread (256 threads)
__device__ __forceinline__ void load_taps(const float* __restrict__ wdw, float (&tv)[CM_TAPLD], int c0, int d, int k) {
  const int nch = min(CM_CH, d - c0);
#pragma unroll
  for (int i = 0; i < CM_TAPLD; ++i) {
    const int e = threadIdx.x + i * 256;
    tv[i] = e < nch * k ? __ldg(wdw + static_cast<int64_t>(c0) * k + e) : 0.f;
  }
}
__device__ __forceinline__ void store_taps(const float (&tv)[CM_TAPLD], float* __restrict__ sWt, int k) {
#pragma unroll
  for (int i = 0; i < CM_TAPLD; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < CM_CH * k) sWt[(e % k) * CM_WP + e / k] = tv[i];
  }
}
__device__ __forceinline__ void stage_taps(const float* __restrict__ wdw, float* __restrict__ sWt, int c0, int d, int k) {
  float tv[CM_TAPLD];
  load_taps(wdw, tv, c0, d, k);
  store_taps(tv, sWt, k);
}

// MODE 0: eval  -> out = SiLU(BN_running(y))            (bn_scale/bn_shift folded per channel)
// MODE 1: train -> y (bf16) + per-channel sum / sumsq   (stats fp64 [2][d], atomics once per CTA per channel)
// KT = compile-time tap count (31, 33) or 0 = run-time k: with KT known the window loop has no branches and no dead taps
template <int MODE, int KT>
__global__ void __launch_bounds__(256, 3) convmod_fwd_kernel(const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ wdw, const float* __restrict__ bdw,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          const float* __restrict__ bn_mean,
                                                          const float* __restrict__ bn_var, float eps,
                                                          __nv_bfloat16* __restrict__ out, double* __restrict__ stats,
                                                          int T, int d, int k_rt) {
  pdl_trigger();
  pdl_wait();
  const int k = KT ? KT : k_rt;
  constexpr int NTAP = KT ? KT : CM_MAXK;
  extern __shared__ float2 sG[];                 // [(CM_TT + k - 1)][32] float2, then taps [k][64] float
  __shared__ float red[2][8][CM_CH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t0 = blockIdx.x * CM_TT, c0 = blockIdx.y * CM_CH, b = blockIdx.z;
  const int half = (k - 1) / 2;
  const int ch = c0 + lane * 2;
  const bool ch_ok = ch < d;
  float* sWt = reinterpret_cast<float*>(sG + (CM_TT + k - 1) * 32);

  {
    float tv[CM_TAPLD];
    load_taps(wdw, tv, c0, d, k);                  // in flight together with the GLU loads below: one memory round trip
    stage_glu(u, sG, b, T, d, c0, t0 - half, CM_TT + k - 1);
    store_taps(tv, sWt, k);
  }
  __syncthreads();

  // taps of this lane's channel pair are fetched from shared memory just in time (tap j is live for window rows
  // j .. j + CM_GF - 1 only): 16 taps in registers instead of all 31 -> 3 CTAs per SM instead of 2
  float2 w[NTAP];
  float2 bias = make_float2(0.f, 0.f);
  if (ch_ok) bias = *reinterpret_cast<const float2*>(bdw + ch);
  float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
  if (MODE == 0 && ch_ok) {
    sc0 = __ldg(bn_w + ch) * rsqrtf(__ldg(bn_var + ch) + eps);
    sc1 = __ldg(bn_w + ch + 1) * rsqrtf(__ldg(bn_var + ch + 1) + eps);
    sh0 = __ldg(bn_b + ch) - __ldg(bn_mean + ch) * sc0;
    sh1 = __ldg(bn_b + ch + 1) - __ldg(bn_mean + ch + 1) * sc1;
  }

  float s_sum0 = 0.f, s_sum1 = 0.f, s_sq0 = 0.f, s_sq1 = 0.f;
  static_assert(CM_TT / 8 == CM_GF, "one register window per warp");
  {
    const int tl = warp * CM_GF;                                // first local output frame of this warp
    float2 a[CM_GF];
#pragma unroll
    for (int i = 0; i < CM_GF; ++i) a[i] = bias;
    // window rows tl .. tl + CM_GF + k - 2 ; output i uses rows tl+i .. tl+i+k-1 with taps 0..k-1
#pragma unroll
    for (int r = 0; r < CM_GF + NTAP - 1; ++r) {
      if (r < NTAP) w[r] = r < k ? *reinterpret_cast<const float2*>(sWt + r * CM_WP + lane * 2) : make_float2(0.f, 0.f);
      if (KT || r < CM_GF + k - 1) {
        const float2 g = sG[(tl + r) * 32 + lane];
#pragma unroll
        for (int i = 0; i < CM_GF; ++i) {
          const int j = r - i;
          if (j >= 0 && j < NTAP) a[i] = ffma2(g, w[j], a[i]);
        }
      }
    }
    __nv_bfloat16* orow = out + (static_cast<int64_t>(b) * T + t0 + tl) * d + ch;
#pragma unroll
    for (int i = 0; i < CM_GF; ++i) {
      if (t0 + tl + i < T && ch_ok) {
        float y0 = a[i].x, y1 = a[i].y;
        if (MODE == 1) {
          // statistics of the STORED (bf16-rounded) tensor: BatchNorm and its backward then see exactly the y they
          // normalise, so sum(yhat) = 0 and the analytically-zero gradients (depthwise bias) stay ~0
          const uint32_t pk = pack_bf16(y0, y1);
          const float2 yr = unpack_bf16(pk);
          s_sum0 += yr.x; s_sum1 += yr.y; s_sq0 = fmaf(yr.x, yr.x, s_sq0); s_sq1 = fmaf(yr.y, yr.y, s_sq1);
          *reinterpret_cast<uint32_t*>(orow + static_cast<int64_t>(i) * d) = pk;
        } else {
          *reinterpret_cast<uint32_t*>(orow + static_cast<int64_t>(i) * d) =
              pack_bf16(silu_f(fmaf(y0, sc0, sh0)), silu_f(fmaf(y1, sc1, sh1)));
        }
      }
    }
  }
  if (MODE == 1) {
    red[0][warp][lane * 2] = s_sum0; red[0][warp][lane * 2 + 1] = s_sum1;
    red[1][warp][lane * 2] = s_sq0;  red[1][warp][lane * 2 + 1] = s_sq1;
    __syncthreads();
    if (threadIdx.x < 2 * CM_CH) {
      const int which = threadIdx.x / CM_CH, c = threadIdx.x % CM_CH;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[which][w][c];
      if (c0 + c < d) atomicAdd(stats + static_cast<int64_t>(which) * d + c0 + c, static_cast<double>(s));
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int d) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const double mu = stats[c] / count;
  double var = stats[d + c] / count - mu * mu;      // biased variance normalises (torch BatchNorm1d, train mode)
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(mu);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// Thread = one fixed group of 8 channels (folded scale/shift live in registers) walking down the rows: one 16-byte load
// and store per 8 elements, four rows in flight per thread.
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ bn_w,
                                                      const float* __restrict__ bn_b, __nv_bfloat16* __restrict__ out,
                                                      int64_t rows, int d, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int cg = d >> 3, rpi = 256 / cg;                 // channel groups per row, rows per CTA step
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  if (tr >= rpi) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j4 = 0; j4 < 2; ++j4) {
    const int c = tg * 8 + j4 * 4;
    const float4 q = __ldg(reinterpret_cast<const float4*>(rstd + c)), gw = __ldg(reinterpret_cast<const float4*>(bn_w + c));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(bn_b + c)), mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    sc[j4 * 4 + 0] = q.x * gw.x; sc[j4 * 4 + 1] = q.y * gw.y; sc[j4 * 4 + 2] = q.z * gw.z; sc[j4 * 4 + 3] = q.w * gw.w;
    sh[j4 * 4 + 0] = gb.x - mu.x * sc[j4 * 4 + 0]; sh[j4 * 4 + 1] = gb.y - mu.y * sc[j4 * 4 + 1];
    sh[j4 * 4 + 2] = gb.z - mu.z * sc[j4 * 4 + 2]; sh[j4 * 4 + 3] = gb.w - mu.w * sc[j4 * 4 + 3];
  }
  const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
  constexpr int UN = 4;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r0 < r_end; r0 += static_cast<int64_t>(UN) * rpi) {
    uint4 v[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int64_t r = r0 + static_cast<int64_t>(q) * rpi;
      if (r < r_end) v[q] = *reinterpret_cast<const uint4*>(y + r * d + tg * 8);
    }
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int64_t r = r0 + static_cast<int64_t>(q) * rpi;
      if (r >= r_end) break;
      const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(w[j]);
        o[j] = pack_bf16(silu_f(fmaf(f.x, sc[2 * j], sh[2 * j])), silu_f(fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1])));
      }
      *reinterpret_cast<uint4*>(out + r * d + tg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// bn_finalize + bn_apply in ONE launch (training forward: dw -> finalize -> apply was three launches; the finalize kernel is two
// CTAs of pure latency between two memory passes).  Every thread derives mean / rstd of its own 8 channels from the fp64
// sums -- the same expressions, rounded to fp32 at the same points, as bn_finalize_kernel, so the result is bit-identical to
// the two-launch path -- and CTA 0 also publishes mean / rstd (the backward pass needs them) and updates the running statistics.
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(const __nv_bfloat16* __restrict__ y, const double* __restrict__ stats,
                                                               double count, float eps, float momentum,
                                                               const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                               __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, float* __restrict__ running_mean,
                                                               float* __restrict__ running_var, int64_t rows, int d,
                                                               int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  const int cg = d >> 3, rpi = 256 / cg;
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  if (tr >= rpi) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = tg * 8 + j;
    const double mu = stats[c] / count;
    double var = stats[d + c] / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mean_f = static_cast<float>(mu);
    const float rstd_f = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    sc[j] = rstd_f * __ldg(bn_w + c);
    sh[j] = __ldg(bn_b + c) - mean_f * sc[j];
    if (blockIdx.x == 0 && tr == 0) {
      mean_out[c] = mean_f;
      rstd_out[c] = rstd_f;
      if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
      }
    }
  }
  const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
  constexpr int UN = 4;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r0 < r_end; r0 += static_cast<int64_t>(UN) * rpi) {
    uint4 v[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int64_t r = r0 + static_cast<int64_t>(q) * rpi;
      if (r < r_end) v[q] = *reinterpret_cast<const uint4*>(y + r * d + tg * 8);
    }
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int64_t r = r0 + static_cast<int64_t>(q) * rpi;
      if (r >= r_end) break;
      const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(w[j]);
        o[j] = pack_bf16(silu_f(fmaf(f.x, sc[2 * j], sh[2 * j])), silu_f(fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1])));
      }
      *reinterpret_cast<uint4*>(out + r * d + tg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}


// ================================================================================================
// backward of the chain
//   dz = dc * SiLU'(z),  z = yhat*gamma + beta,  yhat = (y - mean) * rstd
//   train BN: dy = gamma*rstd * (dz - mean_N(dz) - yhat * mean_N(dz*yhat))     eval BN: dy = gamma*rstd*dz
//   dg[t] = sum_j w[j] * dy[t + h - j]      dw[j] = sum_t dy[t] * g[t + j - h]      dbdw = sum_t dy[t]
//   GLU: da = dg * sigmoid(b),  db = dg * a * sigmoid(b) * (1 - sigmoid(b))
// ================================================================================================

// pass 1: S1 = sum dz, S2 = sum dz*yhat per channel (fp64 atomics, once per CTA per channel) = dbeta, dgamma
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dc,
                                                           const __nv_bfloat16* __restrict__ y,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                           double* __restrict__ sums, int64_t rows, int d,
                                                           int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  // thread = fixed group of 8 channels walking down this CTA's rows (16-byte loads, 2 x 4 rows in flight); the CTA's
  // partial sums meet in shared memory and leave as one fp64 atomic per channel
  extern __shared__ float red_dyn[];                     // [2][d]
  const int cg = d >> 3, rpi = 256 / cg;
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  for (int i = threadIdx.x; i < 2 * d; i += 256) red_dyn[i] = 0.f;
  __syncthreads();
  if (tr < rpi) {
    // per channel only two folded constants stay in registers (z = y*A + C with A = rstd*gamma, C = beta - mean*A) and
    // the sums are taken over dz and dz*y; sum(dz*yhat) = rstd * (sum(dz*y) - mean*sum(dz)) is formed once at the end:
    // ~64 registers -> 4 CTAs per SM (the first version held 6 arrays of 8 and ran at 2 CTAs per SM, issue bound)
    float A[8], C[8], s1[8], s3[8];
#pragma unroll
    for (int j4 = 0; j4 < 2; ++j4) {
      const int c = tg * 8 + j4 * 4;
      const float4 q = __ldg(reinterpret_cast<const float4*>(rstd + c)), gw = __ldg(reinterpret_cast<const float4*>(bn_w + c));
      const float4 gb = __ldg(reinterpret_cast<const float4*>(bn_b + c)), mu = __ldg(reinterpret_cast<const float4*>(mean + c));
      A[j4 * 4 + 0] = q.x * gw.x; A[j4 * 4 + 1] = q.y * gw.y; A[j4 * 4 + 2] = q.z * gw.z; A[j4 * 4 + 3] = q.w * gw.w;
      C[j4 * 4 + 0] = gb.x - mu.x * A[j4 * 4 + 0]; C[j4 * 4 + 1] = gb.y - mu.y * A[j4 * 4 + 1];
      C[j4 * 4 + 2] = gb.z - mu.z * A[j4 * 4 + 2]; C[j4 * 4 + 3] = gb.w - mu.w * A[j4 * 4 + 3];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s3[j] = 0.f;
    const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
    constexpr int UN = 5;                              // ~10 rows per thread: two batches of loads, two memory round trips
    for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r0 < r_end; r0 += static_cast<int64_t>(UN) * rpi) {
      uint4 vd[UN], vy[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t r = min(r0 + static_cast<int64_t>(u) * rpi, r_end - 1);          // clamped: loads stay unconditional
        vd[u] = *reinterpret_cast<const uint4*>(dc + r * d + tg * 8);
        vy[u] = *reinterpret_cast<const uint4*>(y + r * d + tg * 8);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const float live = r0 + static_cast<int64_t>(u) * rpi < r_end ? 1.f : 0.f;
        const uint32_t wd[4] = {vd[u].x, vd[u].y, vd[u].z, vd[u].w}, wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 dcv = unpack_bf16(wd[j]), yv = unpack_bf16(wy[j]);
          const float dz0 = live * dcv.x * silu_grad_f(fmaf(yv.x, A[2 * j], C[2 * j]));
          const float dz1 = live * dcv.y * silu_grad_f(fmaf(yv.y, A[2 * j + 1], C[2 * j + 1]));
          s1[2 * j] += dz0; s1[2 * j + 1] += dz1;
          s3[2 * j] = fmaf(dz0, yv.x, s3[2 * j]); s3[2 * j + 1] = fmaf(dz1, yv.y, s3[2 * j + 1]);
        }
      }
    }
#pragma unroll
    for (int j4 = 0; j4 < 2; ++j4) {
      const int c = tg * 8 + j4 * 4;
      const float4 q = __ldg(reinterpret_cast<const float4*>(rstd + c)), mu = __ldg(reinterpret_cast<const float4*>(mean + c));
      const float qq[4] = {q.x, q.y, q.z, q.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j4 * 4 + e;
        atomicAdd(&red_dyn[c + e], s1[j]);
        atomicAdd(&red_dyn[d + c + e], qq[e] * (s3[j] - mm[e] * s1[j]));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += 256) atomicAdd(sums + i, static_cast<double>(red_dyn[i]));
}

// pass 2: dy on the fly -> du (through depthwise conv + GLU), dwdw, dbdw (fp32 atomics), dbn_w/dbn_b (from sums)
// KT = compile-time tap count (31) or 0 = run-time k (loops sized for CM_MAXK, dead taps multiplied by zero)
template <bool BN_TRAIN, int KT>
__global__ void __launch_bounds__(256, 2) convmod_bwd_kernel(const __nv_bfloat16* __restrict__ dc,
                                                          const __nv_bfloat16* __restrict__ y,
                                                          const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          const double* __restrict__ sums, double inv_count,
                                                          const float* __restrict__ wdw, __nv_bfloat16* __restrict__ du,
                                                          float* __restrict__ dwdw, float* __restrict__ dbdw,
                                                          float* __restrict__ dbn_w, float* __restrict__ dbn_b,
                                                          float* __restrict__ dbias_u, int T, int d, int k_rt) {
  pdl_trigger();
  pdl_wait();
  const int k = KT ? KT : k_rt;
  constexpr int NT = KT ? KT : CM_MAXK;
  extern __shared__ float2 sbuf[];
  const int nrow = CM_TT + k - 1;
  float2* sG = sbuf;                                 // GLU(u)  [nrow][32]
  float2* sDY = sbuf + nrow * 32;                    // dy      [nrow][32]
  float* sW = reinterpret_cast<float*>(sbuf + 2 * nrow * 32);   // dw partials [64][CM_MAXK] + dbdw [64]
  float* sWt = sW + CM_CH * (CM_MAXK + 1);                       // taps [k][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t0 = blockIdx.x * CM_TT, c0 = blockIdx.y * CM_CH, b = blockIdx.z;
  const int half = (k - 1) / 2;
  const int ch = c0 + lane * 2;
  const bool ch_ok = ch < d;

  float tv[CM_TAPLD];
  uint2 av[CM_UN], gv[CM_UN];
  load_taps(wdw, tv, c0, d, k);                      // taps + GLU operands: one batch of global loads; dc / y: a second
  const uint32_t glu_ok = glu_load(u, av, gv, b, T, d, c0, t0 - half, nrow);
  glu_store(av, gv, glu_ok, sG, nrow);
  store_taps(tv, sWt, k);
  {
    if (ch_ok && blockIdx.x == 0 && blockIdx.z == 0 && warp == 0) {     // parameter grads of BN = the two sums
      atomicAdd(dbn_b + ch, static_cast<float>(sums[ch])); atomicAdd(dbn_b + ch + 1, static_cast<float>(sums[ch + 1]));
      atomicAdd(dbn_w + ch, static_cast<float>(sums[d + ch])); atomicAdd(dbn_w + ch + 1, static_cast<float>(sums[d + ch + 1]));
    }
    // same half-warp-per-row / 4-channels-per-lane scheme as stage_glu (the lane's channel quad differs from its
    // compute pair, so the per-channel constants are re-read here through the read-only path)
    const int hl = lane & 15, rsel = lane >> 4;
    const int ch4 = c0 + hl * 4;
    const bool ch4_ok = ch4 < d;
    float pm[4], pq[4], pg[4], pb[4], pa1[4], pa2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pm[j] = pq[j] = pg[j] = pb[j] = pa1[j] = pa2[j] = 0.f;
      if (ch4_ok) {
        pm[j] = __ldg(mean + ch4 + j); pq[j] = __ldg(rstd + ch4 + j); pg[j] = __ldg(bn_w + ch4 + j); pb[j] = __ldg(bn_b + ch4 + j);
        if (BN_TRAIN) {
          pa1[j] = static_cast<float>(sums[ch4 + j] * inv_count);
          pa2[j] = static_cast<float>(sums[d + ch4 + j] * inv_count);
        }
      }
    }
    float4* sDY4 = reinterpret_cast<float4*>(sDY);
    constexpr int UN = CM_UN;
    {
      const int r0 = warp * 2 + rsel;
      uint2 dcr[UN], yr[UN];
      bool ok[UN];
      const int chc = ch4_ok ? ch4 : c0;             // unconditional loads from a clamped address (see glu_load)
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int r = r0 + q * 16, t = t0 - half + r;
        ok[q] = r < nrow && t >= 0 && t < T && ch4_ok;
        const int64_t row = static_cast<int64_t>(b) * T + min(max(t, 0), T - 1);
        dcr[q] = *reinterpret_cast<const uint2*>(dc + row * d + chc);
        yr[q] = *reinterpret_cast<const uint2*>(y + row * d + chc);
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int r = r0 + q * 16;
        if (r >= nrow) continue;
        float dyv[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok[q]) {
          const float2 d01 = unpack_bf16(dcr[q].x), d23 = unpack_bf16(dcr[q].y);
          const float2 y01 = unpack_bf16(yr[q].x), y23 = unpack_bf16(yr[q].y);
          const float dcv[4] = {d01.x, d01.y, d23.x, d23.y}, yv[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float yh = (yv[j] - pm[j]) * pq[j];
            const float dz = dcv[j] * silu_grad_f(fmaf(yh, pg[j], pb[j]));
            dyv[j] = pg[j] * pq[j] * (BN_TRAIN ? (dz - pa1[j] - yh * pa2[j]) : dz);
          }
        }
        sDY4[r * 16 + hl] = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
      }
    }
  }
  __syncthreads();

  float2 dua = make_float2(0.f, 0.f), dug = make_float2(0.f, 0.f);
  // ---- phase 1: dg = correlate(dy, reversed taps) for this warp's 16 frames, then du through the GLU ----
  {
    float2 w[NT];                                     // REVERSED taps: wr[j'] = w[k-1-j']
#pragma unroll
    for (int j = 0; j < NT; ++j)
      w[j] = j < k ? *reinterpret_cast<const float2*>(sWt + (k - 1 - j) * CM_WP + lane * 2) : make_float2(0.f, 0.f);
#pragma unroll 1
    for (int grp = 0; grp < (CM_TT / 8) / CM_G; ++grp) {
      const int tl = warp * (CM_TT / 8) + grp * CM_G;
      // the GLU operands of this group's 8 frames are fetched (unconditionally, clamped) BEFORE the tap window runs, so
      // their latency hides under its 250 FFMA2s instead of being exposed once per frame in the store loop below
      uint32_t ua[CM_G], ug[CM_G];
      {
        const __nv_bfloat16* ub = u + static_cast<int64_t>(b) * T * (2 * d) + (ch_ok ? ch : c0);
#pragma unroll
        for (int i = 0; i < CM_G; ++i) {
          const __nv_bfloat16* rowp = ub + static_cast<int64_t>(min(t0 + tl + i, T - 1)) * (2 * d);
          ua[i] = *reinterpret_cast<const uint32_t*>(rowp);
          ug[i] = *reinterpret_cast<const uint32_t*>(rowp + d);
        }
      }
      float2 a[CM_G];
#pragma unroll
      for (int i = 0; i < CM_G; ++i) a[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < CM_G + NT - 1; ++r) {
        if (KT || r < CM_G + k - 1) {
          const float2 dyr = sDY[(tl + r) * 32 + lane];
#pragma unroll
          for (int i = 0; i < CM_G; ++i) {
            const int j = r - i;
            if (j >= 0 && j < NT) a[i] = ffma2(dyr, w[j], a[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CM_G; ++i) {
        const int t = t0 + tl + i;
        if (t < T && ch_ok) {
          const int64_t row = static_cast<int64_t>(b) * T + t;
          const float2 av = unpack_bf16(ua[i]), bv = unpack_bf16(ug[i]);
          const float s0 = sigmoid_f(bv.x), s1 = sigmoid_f(bv.y);
          const float da0 = a[i].x * s0, da1 = a[i].y * s1;
          const float dg0 = a[i].x * av.x * s0 * (1.f - s0), dg1 = a[i].y * av.y * s1 * (1.f - s1);
          *reinterpret_cast<uint32_t*>(du + row * 2 * d + ch) = pack_bf16(da0, da1);
          *reinterpret_cast<uint32_t*>(du + row * 2 * d + d + ch) = pack_bf16(dg0, dg1);
          dua.x += da0; dua.y += da1; dug.x += dg0; dug.y += dg1;       // bias gradient of the pointwise conv that made u
        }
      }
    }
  }

  // ---- phase 2: dw[j] += dy[t] * g[t + j - h] over this warp's frames (taps are dead: registers are reused) ----
  float2 acc[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j] = make_float2(0.f, 0.f);
  float dbs0 = 0.f, dbs1 = 0.f;
#pragma unroll 1
  for (int grp = 0; grp < (CM_TT / 8) / CM_G; ++grp) {
    const int tl = warp * (CM_TT / 8) + grp * CM_G;
    float2 dyc[CM_G];                                 // dy at the 8 centre frames
#pragma unroll
    for (int i = 0; i < CM_G; ++i) {
      dyc[i] = sDY[(tl + half + i) * 32 + lane];
      dbs0 += dyc[i].x; dbs1 += dyc[i].y;
    }
#pragma unroll
    for (int r = 0; r < CM_G + NT - 1; ++r) {
      if (KT || r < CM_G + k - 1) {
        const float2 gr = sG[(tl + r) * 32 + lane];
#pragma unroll
        for (int i = 0; i < CM_G; ++i) {
          const int j = r - i;
          if (j >= 0 && j < NT) acc[j] = ffma2(dyc[i], gr, acc[j]);
        }
      }
    }
  }
  // CTA-level reduction of dw / dbdw over the 8 warps: the staged tiles are dead now, so every warp parks its partials
  // in their place ([warp][tap][64 ch], conflict-free float2 stores -- the first version used 66 shared-memory
  // atomics per thread here), 256 threads add the 8 copies, then one global atomic per (channel, tap)
  __syncthreads();
  float2* red2 = sbuf;                               // [8][CM_RS][32] float2  (73.7 KB <= the staged tiles + the old dw area)
#pragma unroll
  for (int j = 0; j < NT; ++j) red2[(warp * CM_RS + j) * 32 + lane] = acc[j];
  red2[(warp * CM_RS + CM_MAXK) * 32 + lane] = make_float2(dbs0, dbs1);
  red2[(warp * CM_RS + CM_MAXK + 1) * 32 + lane] = dua;
  red2[(warp * CM_RS + CM_MAXK + 2) * 32 + lane] = dug;
  __syncthreads();
  const float* red = reinterpret_cast<const float*>(sbuf);
  for (int e = threadIdx.x; e < CM_RS * CM_CH; e += blockDim.x) {
    const int j = e / CM_CH, c = e % CM_CH;
    if (c0 + c >= d || (j >= k && j < CM_MAXK) || (j > CM_MAXK && dbias_u == nullptr)) continue;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[(w * CM_RS + j) * CM_CH + c];
    if (j == CM_MAXK) atomicAdd(dbdw + c0 + c, sum);
    else if (j == CM_MAXK + 1) atomicAdd(dbias_u + c0 + c, sum);
    else if (j == CM_MAXK + 2) atomicAdd(dbias_u + d + c0 + c, sum);
    else atomicAdd(dwdw + static_cast<int64_t>(c0 + c) * k + j, sum);
  }
}

static int convmod_fwd_attr() {
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<0, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<1, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<0, 33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_fwd_kernel<1, 33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  return 0;
}

static int check_convmod(int B, int T, int d, int k, const char* who) {
  B2C_REQUIRE(B > 0 && T > 0 && d > 0, "%s: empty problem", who);
  B2C_REQUIRE(d % 8 == 0, "%s: d=%d must be a multiple of 8", who, d);
  B2C_REQUIRE(k % 2 == 1 && k >= 1 && k <= CM_MAXK, "%s: depthwise kernel %d must be odd and <= %d", who, k, CM_MAXK);
  return 0;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_convmod_eval_fwd(const void* u, const float* wdw, const float* bdw, const float* bn_w,
                                    const float* bn_b, const float* bn_mean, const float* bn_var, float eps, void* out,
                                    int B, int T, int d, int k, void* stream) {
  B2C_REQUIRE(u && wdw && bdw && bn_w && bn_b && bn_mean && bn_var && out, "b2c_convmod_eval_fwd: null pointer");
  if (int rc = check_convmod(B, T, d, k, "b2c_convmod_eval_fwd")) return rc;
  dim3 grid((T + CM_TT - 1) / CM_TT, (d + CM_CH - 1) / CM_CH, B);
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) + CM_WP * k * sizeof(float);
  if (int rc = convmod_fwd_attr()) return rc;
#define B2C_CM_FWD(KT_)                                                                                          \
  B2C_CUDA(launch_k(convmod_fwd_kernel<0, KT_>, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream),                            \
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, bn_w, bn_b, bn_mean, bn_var, eps, static_cast<__nv_bfloat16*>(out), \
      nullptr, T, d, k))
  if (k == 31) B2C_CM_FWD(31); else if (k == 33) B2C_CM_FWD(33); else B2C_CM_FWD(0);
#undef B2C_CM_FWD
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_dw_fwd(const void* u, const float* wdw, const float* bdw, void* y, double* stats, int B,
                                  int T, int d, int k, void* stream) {
  B2C_REQUIRE(u && wdw && bdw && y && stats, "b2c_convmod_dw_fwd: null pointer");
  if (int rc = check_convmod(B, T, d, k, "b2c_convmod_dw_fwd")) return rc;
  dim3 grid((T + CM_TT - 1) / CM_TT, (d + CM_CH - 1) / CM_CH, B);
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) + CM_WP * k * sizeof(float);
  if (int rc = convmod_fwd_attr()) return rc;
#define B2C_CM_FWD(KT_)                                                                                          \
  B2C_CUDA(launch_k(convmod_fwd_kernel<1, KT_>, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream),                            \
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, nullptr, nullptr, nullptr, nullptr, 0.f,                   \
      static_cast<__nv_bfloat16*>(y), stats, T, d, k))
  if (k == 31) B2C_CM_FWD(31); else if (k == 33) B2C_CM_FWD(33); else B2C_CM_FWD(0);
#undef B2C_CM_FWD
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_finalize(const double* stats, int64_t count, float eps, float momentum, float* mean,
                                       float* rstd, float* running_mean, float* running_var, int d, void* stream) {
  B2C_REQUIRE(stats && mean && rstd && d > 0 && count > 0, "b2c_convmod_bn_finalize: bad argument");
  B2C_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "b2c_convmod_bn_finalize: running stats mismatch");
  B2C_CUDA(launch_k(bn_finalize_kernel, dim3((d + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      stats, static_cast<double>(count), eps, momentum, mean, rstd, running_mean, running_var, d));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_apply(const void* y, const float* mean, const float* rstd, const float* bn_w,
                                    const float* bn_b, void* out, int64_t rows, int d, void* stream) {
  B2C_REQUIRE(y && mean && rstd && bn_w && bn_b && out, "b2c_convmod_bn_apply: null pointer");
  B2C_REQUIRE(d % 8 == 0, "b2c_convmod_bn_apply: d must be a multiple of 8");
  B2C_REQUIRE(d <= 2048, "b2c_convmod_bn_apply: d must be <= 2048");
  if (rows <= 0) return 0;
  const int rpi = 256 / (d / 8);
  int64_t nblk = static_cast<int64_t>(num_sms()) * 4;    // a few fat CTAs: the per-thread constants are set up once
  int64_t rpb = (rows + nblk - 1) / nblk;
  rpb = ((rpb + rpi - 1) / rpi) * rpi;
  nblk = (rows + rpb - 1) / rpb;
  B2C_CUDA(launch_k(bn_apply_kernel, dim3(static_cast<unsigned>(nblk)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(y), mean, rstd, bn_w, bn_b, static_cast<__nv_bfloat16*>(out), rows, d,
      static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_finalize_apply(const void* y, const double* stats, int64_t count, float eps, float momentum,
                                            const float* bn_w, const float* bn_b, void* out, float* mean, float* rstd,
                                            float* running_mean, float* running_var, int64_t rows, int d, void* stream) {
  B2C_REQUIRE(y && stats && bn_w && bn_b && out && mean && rstd && count > 0, "b2c_convmod_bn_finalize_apply: bad argument");
  B2C_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "b2c_convmod_bn_finalize_apply: running stats mismatch");
  B2C_REQUIRE(d % 8 == 0 && d <= 2048, "b2c_convmod_bn_finalize_apply: d must be a multiple of 8, <= 2048");
  if (rows <= 0) return 0;
  const int rpi = 256 / (d / 8);
  int64_t nblk = static_cast<int64_t>(num_sms()) * 4;
  int64_t rpb = (rows + nblk - 1) / nblk;
  rpb = ((rpb + rpi - 1) / rpi) * rpi;
  nblk = (rows + rpb - 1) / rpb;
  B2C_CUDA(launch_k(bn_finalize_apply_kernel, dim3(static_cast<unsigned>(nblk)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      static_cast<const __nv_bfloat16*>(y), stats, static_cast<double>(count), eps, momentum, bn_w, bn_b,
      static_cast<__nv_bfloat16*>(out), mean, rstd, running_mean, running_var, rows, d, static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_bwd_reduce(const void* dc, const void* y, const float* mean, const float* rstd,
                                         const float* bn_w, const float* bn_b, double* sums, int64_t rows, int d,
                                         void* stream) {
  B2C_REQUIRE(dc && y && mean && rstd && bn_w && bn_b && sums, "b2c_convmod_bn_bwd_reduce: null pointer");
  B2C_REQUIRE(d % 8 == 0, "b2c_convmod_bn_bwd_reduce: d must be a multiple of 8");
  if (rows <= 0) return 0;
  B2C_REQUIRE(d <= 2048, "b2c_convmod_bn_bwd_reduce: d must be <= 2048");
  const int rpi = 256 / (d / 8);
  static const int per_sm = [] { const char* e = getenv("B2C_BN_RED_PER_SM"); return e ? atoi(e) : 1; }();
  int64_t nblk = static_cast<int64_t>(num_sms()) * per_sm;   // 1 CTA per SM by default: every CTA ends with 2*d fp64 atomics and those dominate (12 / 14 / 18 us at 1 / 2 / 4 per SM)
  int64_t rpb = (rows + nblk - 1) / nblk;
  rpb = ((rpb + rpi - 1) / rpi) * rpi;
  nblk = (rows + rpb - 1) / rpb;
  B2C_CUDA(launch_k(bn_bwd_reduce_kernel, dim3(static_cast<unsigned>(nblk)), dim3(256), 2 * d * sizeof(float), static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(dc), static_cast<const __nv_bfloat16*>(y), mean, rstd, bn_w, bn_b, sums, rows, d,
      static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_dw_bwd(const void* dc, const void* y, const void* u, const float* mean, const float* rstd,
                                  const float* bn_w, const float* bn_b, const double* sums, int bn_train,
                                  const float* wdw, void* du, float* dwdw, float* dbdw, float* dbn_w, float* dbn_b,
                                  float* dbias_u, int B, int T, int d, int k, void* stream) {
  B2C_REQUIRE(dc && y && u && mean && rstd && bn_w && bn_b && sums && wdw && du && dwdw && dbdw && dbn_w && dbn_b,
              "b2c_convmod_dw_bwd: null pointer");
  if (int rc = check_convmod(B, T, d, k, "b2c_convmod_dw_bwd")) return rc;
  dim3 grid((T + CM_TT - 1) / CM_TT, (d + CM_CH - 1) / CM_CH, B);
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) * 2 + CM_CH * (CM_MAXK + 1) * sizeof(float) +
                      CM_WP * k * sizeof(float);
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<true, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<false, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<true, 33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<false, 33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  const double inv_count = 1.0 / (static_cast<double>(B) * T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto args = [&](auto kern) -> int {
    B2C_CUDA(launch_k(kern, dim3(grid), dim3(256), smem, st, static_cast<const __nv_bfloat16*>(dc), static_cast<const __nv_bfloat16*>(y),
                                  static_cast<const __nv_bfloat16*>(u), mean, rstd, bn_w, bn_b, sums, inv_count, wdw,
                                  static_cast<__nv_bfloat16*>(du), dwdw, dbdw, dbn_w, dbn_b, dbias_u, T, d, k));
    return 0;
  };
  int rc;
  if (k == 31)      rc = bn_train ? args(convmod_bwd_kernel<true, 31>) : args(convmod_bwd_kernel<false, 31>);
  else if (k == 33) rc = bn_train ? args(convmod_bwd_kernel<true, 33>) : args(convmod_bwd_kernel<false, 33>);
  else              rc = bn_train ? args(convmod_bwd_kernel<true, 0>) : args(convmod_bwd_kernel<false, 0>);
  if (rc) return rc;
  B2C_CUDA(cudaGetLastError());
  return 0;
}
