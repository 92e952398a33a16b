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
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int CM_TT = 128;     // output frames per CTA
constexpr int CM_CH = 64;      // channels per CTA
constexpr int CM_MAXK = 33;    // taps (repo default k=33, Conformer S/M/L k=31)
constexpr int CM_G = 8;        // outputs per register window

// stage GLU(u) for frames [t_lo, t_lo + nrow) of utterance b into smem (zero outside [0,T)).
// Rows are taken 4 at a time per warp with all 8 global loads issued before any is consumed (memory-level
// parallelism: the un-batched version exposed one DRAM round trip per row and dominated the kernel).
__device__ __forceinline__ void stage_glu(const __nv_bfloat16* __restrict__ u, float2* __restrict__ sG, int b, int T,
                                          int d, int c0, int t_lo, int nrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = c0 + lane * 2;
  const bool ch_ok = ch < d;
  constexpr int UN = 4;
  for (int r0 = warp; r0 < nrow; r0 += 8 * UN) {
    uint32_t av[UN], gv[UN];
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int r = r0 + q * 8, t = t_lo + r;
      av[q] = gv[q] = 0u;
      if (r < nrow && t >= 0 && t < T && ch_ok) {
        const __nv_bfloat16* row = u + (static_cast<int64_t>(b) * T + t) * (2 * d);
        av[q] = *reinterpret_cast<const uint32_t*>(row + ch);
        gv[q] = *reinterpret_cast<const uint32_t*>(row + d + ch);
      }
    }
#pragma unroll
    for (int q = 0; q < UN; ++q) {
      const int r = r0 + q * 8;
      if (r < nrow) {
        const float2 a = unpack_bf16(av[q]), gt = unpack_bf16(gv[q]);
        sG[r * 32 + lane] = make_float2(a.x * sigmoid_f(gt.x), a.y * sigmoid_f(gt.y));   // zero rows: 0 * 0.5 = 0
      }
    }
  }
}

// depthwise taps of this CTA's 64 channels -> smem, tap-major [k][64] (coalesced global read, conflict-free float2 reads)
__device__ __forceinline__ void stage_taps(const float* __restrict__ wdw, float* __restrict__ sWt, int c0, int d, int k) {
  const int nch = min(CM_CH, d - c0);
  for (int e = threadIdx.x; e < nch * k; e += blockDim.x) sWt[(e % k) * CM_CH + e / k] = __ldg(wdw + static_cast<int64_t>(c0) * k + e);
  for (int e = nch * k + threadIdx.x; e < CM_CH * k; e += blockDim.x) sWt[(e % k) * CM_CH + e / k] = 0.f;
}

// MODE 0: eval  -> out = SiLU(BN_running(y))            (bn_scale/bn_shift folded per channel)
// MODE 1: train -> y (bf16) + per-channel sum / sumsq   (stats fp64 [2][d], atomics once per CTA per channel)
// KT = compile-time tap count (31, 33) or 0 = run-time k: with KT known the window loop has no branches and no dead taps
template <int MODE, int KT>
__global__ void __launch_bounds__(256) convmod_fwd_kernel(const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ wdw, const float* __restrict__ bdw,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          const float* __restrict__ bn_mean,
                                                          const float* __restrict__ bn_var, float eps,
                                                          __nv_bfloat16* __restrict__ out, double* __restrict__ stats,
                                                          int T, int d, int k_rt) {
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

  stage_taps(wdw, sWt, c0, d, k);
  stage_glu(u, sG, b, T, d, c0, t0 - half, CM_TT + k - 1);
  __syncthreads();

  float w0[NTAP], w1[NTAP];
#pragma unroll
  for (int j = 0; j < NTAP; ++j) {
    const float2 wv = j < k ? *reinterpret_cast<const float2*>(sWt + j * CM_CH + lane * 2) : make_float2(0.f, 0.f);
    w0[j] = wv.x; w1[j] = wv.y;
  }
  const float bias0 = ch_ok ? __ldg(bdw + ch) : 0.f, bias1 = ch_ok ? __ldg(bdw + ch + 1) : 0.f;
  float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
  if (MODE == 0 && ch_ok) {
    sc0 = __ldg(bn_w + ch) * rsqrtf(__ldg(bn_var + ch) + eps);
    sc1 = __ldg(bn_w + ch + 1) * rsqrtf(__ldg(bn_var + ch + 1) + eps);
    sh0 = __ldg(bn_b + ch) - __ldg(bn_mean + ch) * sc0;
    sh1 = __ldg(bn_b + ch + 1) - __ldg(bn_mean + ch + 1) * sc1;
  }

  float s_sum0 = 0.f, s_sum1 = 0.f, s_sq0 = 0.f, s_sq1 = 0.f;
#pragma unroll 1
  for (int grp = 0; grp < (CM_TT / 8) / CM_G; ++grp) {          // 16 frames per warp = 2 groups of 8
    const int tl = warp * (CM_TT / 8) + grp * CM_G;             // first local output frame of the group
    float a0[CM_G], a1[CM_G];
#pragma unroll
    for (int i = 0; i < CM_G; ++i) { a0[i] = bias0; a1[i] = bias1; }
    // window rows tl .. tl + CM_G + k - 2 ; output i uses rows tl+i .. tl+i+k-1 with taps 0..k-1
#pragma unroll
    for (int r = 0; r < CM_G + NTAP - 1; ++r) {
      if (KT || r < CM_G + k - 1) {
        const float2 g = sG[(tl + r) * 32 + lane];
#pragma unroll
        for (int i = 0; i < CM_G; ++i) {
          const int j = r - i;
          if (j >= 0 && j < NTAP) { a0[i] = fmaf(g.x, w0[j], a0[i]); a1[i] = fmaf(g.y, w1[j], a1[i]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CM_G; ++i) {
      const int t = t0 + tl + i;
      if (t < T && ch_ok) {
        float y0 = a0[i], y1 = a1[i];
        if (MODE == 1) {
          // statistics of the STORED (bf16-rounded) tensor: BatchNorm and its backward then see exactly the y they
          // normalise, so sum(yhat) = 0 and the analytically-zero gradients (depthwise bias) stay ~0
          const float2 yr = unpack_bf16(pack_bf16(y0, y1));
          y0 = yr.x; y1 = yr.y;
          s_sum0 += y0; s_sum1 += y1; s_sq0 += y0 * y0; s_sq1 += y1 * y1;
        } else {
          y0 = silu_f(y0 * sc0 + sh0);
          y1 = silu_f(y1 * sc1 + sh1);
        }
        *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * T + t) * d + ch) = pack_bf16(y0, y1);
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

__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ bn_w,
                                                      const float* __restrict__ bn_b, __nv_bfloat16* __restrict__ out,
                                                      int64_t n8, int d) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // index of an 8-element chunk
  if (i >= n8) return;
  const int c = static_cast<int>((i * 8) % d);
  const uint4 v = reinterpret_cast<const uint4*>(y)[i];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16(w[j]);
    const int cc = c + 2 * j;
    const float z0 = (f.x - __ldg(mean + cc)) * __ldg(rstd + cc) * __ldg(bn_w + cc) + __ldg(bn_b + cc);
    const float z1 = (f.y - __ldg(mean + cc + 1)) * __ldg(rstd + cc + 1) * __ldg(bn_w + cc + 1) + __ldg(bn_b + cc + 1);
    o[j] = pack_bf16(silu_f(z0), silu_f(z1));
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}


// ================================================================================================
// backward of the chain
//   dz = dc * SiLU'(z),  z = yhat*gamma + beta,  yhat = (y - mean) * rstd
//   train BN: dy = gamma*rstd * (dz - mean_N(dz) - yhat * mean_N(dz*yhat))     eval BN: dy = gamma*rstd*dz
//   dg[t] = sum_j w[j] * dy[t + h - j]      dw[j] = sum_t dy[t] * g[t + j - h]      dbdw = sum_t dy[t]
//   GLU: da = dg * sigmoid(b),  db = dg * a * sigmoid(b) * (1 - sigmoid(b))
// ================================================================================================

// pass 1: S1 = sum dz, S2 = sum dz*yhat per channel (fp64 atomics, once per CTA per channel) = dbeta, dgamma
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dc,
                                                           const __nv_bfloat16* __restrict__ y,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                           double* __restrict__ sums, int64_t rows, int d,
                                                           int rows_per_block) {
  __shared__ float red[2][8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ch = blockIdx.x * 64 + tx * 2;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
  if (ch < d) {
    const float m0 = mean[ch], m1 = mean[ch + 1], q0 = rstd[ch], q1 = rstd[ch + 1];
    const float g0 = bn_w[ch], g1 = bn_w[ch + 1], b0 = bn_b[ch], b1 = bn_b[ch + 1];
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float2 dcv = unpack_bf16(*reinterpret_cast<const uint32_t*>(dc + r * d + ch));
      const float2 yv = unpack_bf16(*reinterpret_cast<const uint32_t*>(y + r * d + ch));
      const float yh0 = (yv.x - m0) * q0, yh1 = (yv.y - m1) * q1;
      const float dz0 = dcv.x * silu_grad_f(yh0 * g0 + b0), dz1 = dcv.y * silu_grad_f(yh1 * g1 + b1);
      s1a += dz0; s1b += dz1; s2a += dz0 * yh0; s2b += dz1 * yh1;
    }
  }
  red[0][ty][tx * 2] = s1a; red[0][ty][tx * 2 + 1] = s1b;
  red[1][ty][tx * 2] = s2a; red[1][ty][tx * 2 + 1] = s2b;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[which][w][c];
    if (blockIdx.x * 64 + c < d) atomicAdd(sums + static_cast<int64_t>(which) * d + blockIdx.x * 64 + c, static_cast<double>(sum));
  }
}

// pass 2: dy on the fly -> du (through depthwise conv + GLU), dwdw, dbdw (fp32 atomics), dbn_w/dbn_b (from sums)
template <bool BN_TRAIN>
__global__ void __launch_bounds__(256) convmod_bwd_kernel(const __nv_bfloat16* __restrict__ dc,
                                                          const __nv_bfloat16* __restrict__ y,
                                                          const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          const double* __restrict__ sums, double inv_count,
                                                          const float* __restrict__ wdw, __nv_bfloat16* __restrict__ du,
                                                          float* __restrict__ dwdw, float* __restrict__ dbdw,
                                                          float* __restrict__ dbn_w, float* __restrict__ dbn_b, int T,
                                                          int d, int k) {
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
  for (int i = threadIdx.x; i < CM_CH * (CM_MAXK + 1); i += blockDim.x) sW[i] = 0.f;

  stage_taps(wdw, sWt, c0, d, k);
  stage_glu(u, sG, b, T, d, c0, t0 - half, nrow);
  {
    float m0 = 0, m1 = 0, q0 = 0, q1 = 0, g0 = 0, g1 = 0, b0 = 0, b1 = 0, a10 = 0, a11 = 0, a20 = 0, a21 = 0;
    if (ch_ok) {
      m0 = mean[ch]; m1 = mean[ch + 1]; q0 = rstd[ch]; q1 = rstd[ch + 1];
      g0 = bn_w[ch]; g1 = bn_w[ch + 1]; b0 = bn_b[ch]; b1 = bn_b[ch + 1];
      if (BN_TRAIN) {
        a10 = static_cast<float>(sums[ch] * inv_count); a11 = static_cast<float>(sums[ch + 1] * inv_count);
        a20 = static_cast<float>(sums[d + ch] * inv_count); a21 = static_cast<float>(sums[d + ch + 1] * inv_count);
      }
      if (blockIdx.x == 0 && blockIdx.z == 0 && warp == 0) {     // parameter grads of BN = the two sums
        atomicAdd(dbn_b + ch, static_cast<float>(sums[ch])); atomicAdd(dbn_b + ch + 1, static_cast<float>(sums[ch + 1]));
        atomicAdd(dbn_w + ch, static_cast<float>(sums[d + ch])); atomicAdd(dbn_w + ch + 1, static_cast<float>(sums[d + ch + 1]));
      }
    }
    constexpr int UN = 4;
    for (int r0 = warp; r0 < nrow; r0 += 8 * UN) {
      uint32_t dcr[UN], yr[UN];
      bool ok[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int r = r0 + q * 8, t = t0 - half + r;
        ok[q] = r < nrow && t >= 0 && t < T && ch_ok;
        dcr[q] = yr[q] = 0u;
        if (ok[q]) {
          const int64_t row = static_cast<int64_t>(b) * T + t;
          dcr[q] = *reinterpret_cast<const uint32_t*>(dc + row * d + ch);
          yr[q] = *reinterpret_cast<const uint32_t*>(y + row * d + ch);
        }
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int r = r0 + q * 8;
        if (r >= nrow) continue;
        float2 dyv = make_float2(0.f, 0.f);
        if (ok[q]) {
          const float2 dcv = unpack_bf16(dcr[q]), yv = unpack_bf16(yr[q]);
          const float yh0 = (yv.x - m0) * q0, yh1 = (yv.y - m1) * q1;
          const float dz0 = dcv.x * silu_grad_f(yh0 * g0 + b0), dz1 = dcv.y * silu_grad_f(yh1 * g1 + b1);
          dyv.x = g0 * q0 * (BN_TRAIN ? (dz0 - a10 - yh0 * a20) : dz0);
          dyv.y = g1 * q1 * (BN_TRAIN ? (dz1 - a11 - yh1 * a21) : dz1);
        }
        sDY[r * 32 + lane] = dyv;
      }
    }
  }
  __syncthreads();

  // ---- phase 1: dg = correlate(dy, reversed taps) for this warp's 16 frames, then du through the GLU ----
  {
    float w0[CM_MAXK], w1[CM_MAXK];                   // REVERSED taps: wr[j'] = w[k-1-j']
#pragma unroll
    for (int j = 0; j < CM_MAXK; ++j) {
      const float2 wv = j < k ? *reinterpret_cast<const float2*>(sWt + (k - 1 - j) * CM_CH + lane * 2) : make_float2(0.f, 0.f);
      w0[j] = wv.x; w1[j] = wv.y;
    }
#pragma unroll 1
    for (int grp = 0; grp < (CM_TT / 8) / CM_G; ++grp) {
      const int tl = warp * (CM_TT / 8) + grp * CM_G;
      float a0[CM_G], a1[CM_G];
#pragma unroll
      for (int i = 0; i < CM_G; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
#pragma unroll
      for (int r = 0; r < CM_G + CM_MAXK - 1; ++r) {
        if (r < CM_G + k - 1) {
          const float2 dyr = sDY[(tl + r) * 32 + lane];
#pragma unroll
          for (int i = 0; i < CM_G; ++i) {
            const int j = r - i;
            if (j >= 0 && j < CM_MAXK) { a0[i] = fmaf(dyr.x, w0[j], a0[i]); a1[i] = fmaf(dyr.y, w1[j], a1[i]); }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CM_G; ++i) {
        const int t = t0 + tl + i;
        if (t < T && ch_ok) {
          const int64_t row = static_cast<int64_t>(b) * T + t;
          const float2 av = unpack_bf16(*reinterpret_cast<const uint32_t*>(u + row * 2 * d + ch));
          const float2 bv = unpack_bf16(*reinterpret_cast<const uint32_t*>(u + row * 2 * d + d + ch));
          const float s0 = sigmoid_f(bv.x), s1 = sigmoid_f(bv.y);
          *reinterpret_cast<uint32_t*>(du + row * 2 * d + ch) = pack_bf16(a0[i] * s0, a1[i] * s1);
          *reinterpret_cast<uint32_t*>(du + row * 2 * d + d + ch) =
              pack_bf16(a0[i] * av.x * s0 * (1.f - s0), a1[i] * av.y * s1 * (1.f - s1));
        }
      }
    }
  }

  // ---- phase 2: dw[j] += dy[t] * g[t + j - h] over this warp's frames (taps are dead: registers are reused) ----
  float acc0[CM_MAXK], acc1[CM_MAXK];
#pragma unroll
  for (int j = 0; j < CM_MAXK; ++j) acc0[j] = acc1[j] = 0.f;
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
    for (int r = 0; r < CM_G + CM_MAXK - 1; ++r) {
      if (r < CM_G + k - 1) {
        const float2 gr = sG[(tl + r) * 32 + lane];
#pragma unroll
        for (int i = 0; i < CM_G; ++i) {
          const int j = r - i;
          if (j >= 0 && j < CM_MAXK) { acc0[j] = fmaf(dyc[i].x, gr.x, acc0[j]); acc1[j] = fmaf(dyc[i].y, gr.y, acc1[j]); }
        }
      }
    }
  }
  // CTA-level reduction of dw / dbdw over the 8 warps (shared-memory atomics), then one global atomic each
#pragma unroll
  for (int j = 0; j < CM_MAXK; ++j) {
    if (j < k) {
      atomicAdd(&sW[(lane * 2) * (CM_MAXK + 1) + j], acc0[j]);
      atomicAdd(&sW[(lane * 2 + 1) * (CM_MAXK + 1) + j], acc1[j]);
    }
  }
  atomicAdd(&sW[(lane * 2) * (CM_MAXK + 1) + CM_MAXK], dbs0);
  atomicAdd(&sW[(lane * 2 + 1) * (CM_MAXK + 1) + CM_MAXK], dbs1);
  __syncthreads();
  for (int i = threadIdx.x; i < CM_CH * (CM_MAXK + 1); i += blockDim.x) {
    const int c = i / (CM_MAXK + 1), j = i % (CM_MAXK + 1);
    if (c0 + c >= d) continue;
    if (j == CM_MAXK) atomicAdd(dbdw + c0 + c, sW[i]);
    else if (j < k) atomicAdd(dwdw + static_cast<int64_t>(c0 + c) * k + j, sW[i]);
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
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) + CM_CH * k * sizeof(float);
  if (int rc = convmod_fwd_attr()) return rc;
#define B2C_CM_FWD(KT_)                                                                                          \
  convmod_fwd_kernel<0, KT_><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(                           \
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, bn_w, bn_b, bn_mean, bn_var, eps, static_cast<__nv_bfloat16*>(out), \
      nullptr, T, d, k)
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
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) + CM_CH * k * sizeof(float);
  if (int rc = convmod_fwd_attr()) return rc;
#define B2C_CM_FWD(KT_)                                                                                          \
  convmod_fwd_kernel<1, KT_><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(                           \
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, nullptr, nullptr, nullptr, nullptr, 0.f,                   \
      static_cast<__nv_bfloat16*>(y), stats, T, d, k)
  if (k == 31) B2C_CM_FWD(31); else if (k == 33) B2C_CM_FWD(33); else B2C_CM_FWD(0);
#undef B2C_CM_FWD
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_finalize(const double* stats, int64_t count, float eps, float momentum, float* mean,
                                       float* rstd, float* running_mean, float* running_var, int d, void* stream) {
  B2C_REQUIRE(stats && mean && rstd && d > 0 && count > 0, "b2c_convmod_bn_finalize: bad argument");
  B2C_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "b2c_convmod_bn_finalize: running stats mismatch");
  bn_finalize_kernel<<<(d + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, static_cast<double>(count), eps, momentum, mean, rstd, running_mean, running_var, d);
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_apply(const void* y, const float* mean, const float* rstd, const float* bn_w,
                                    const float* bn_b, void* out, int64_t rows, int d, void* stream) {
  B2C_REQUIRE(y && mean && rstd && bn_w && bn_b && out, "b2c_convmod_bn_apply: null pointer");
  B2C_REQUIRE(d % 8 == 0, "b2c_convmod_bn_apply: d must be a multiple of 8");
  const int64_t n8 = rows * d / 8;
  if (n8 <= 0) return 0;
  bn_apply_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), mean, rstd, bn_w, bn_b, static_cast<__nv_bfloat16*>(out), n8, d);
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_bn_bwd_reduce(const void* dc, const void* y, const float* mean, const float* rstd,
                                         const float* bn_w, const float* bn_b, double* sums, int64_t rows, int d,
                                         void* stream) {
  B2C_REQUIRE(dc && y && mean && rstd && bn_w && bn_b && sums, "b2c_convmod_bn_bwd_reduce: null pointer");
  B2C_REQUIRE(d % 8 == 0, "b2c_convmod_bn_bwd_reduce: d must be a multiple of 8");
  if (rows <= 0) return 0;
  const unsigned gx = static_cast<unsigned>((d + 63) / 64);
  int64_t want_y = (static_cast<int64_t>(num_sms()) * 4 + gx - 1) / gx;
  int rows_per_block = static_cast<int>((rows + want_y - 1) / want_y);
  rows_per_block = ((rows_per_block + 7) / 8) * 8;
  const unsigned gy = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  bn_bwd_reduce_kernel<<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dc), static_cast<const __nv_bfloat16*>(y), mean, rstd, bn_w, bn_b, sums, rows, d,
      rows_per_block);
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_dw_bwd(const void* dc, const void* y, const void* u, const float* mean, const float* rstd,
                                  const float* bn_w, const float* bn_b, const double* sums, int bn_train,
                                  const float* wdw, void* du, float* dwdw, float* dbdw, float* dbn_w, float* dbn_b,
                                  int B, int T, int d, int k, void* stream) {
  B2C_REQUIRE(dc && y && u && mean && rstd && bn_w && bn_b && sums && wdw && du && dwdw && dbdw && dbn_w && dbn_b,
              "b2c_convmod_dw_bwd: null pointer");
  if (int rc = check_convmod(B, T, d, k, "b2c_convmod_dw_bwd")) return rc;
  dim3 grid((T + CM_TT - 1) / CM_TT, (d + CM_CH - 1) / CM_CH, B);
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2) * 2 + CM_CH * (CM_MAXK + 1) * sizeof(float) +
                      CM_CH * k * sizeof(float);
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    B2C_CUDA(cudaFuncSetAttribute(convmod_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  const double inv_count = 1.0 / (static_cast<double>(B) * T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto args = [&](auto kern) {
    kern<<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(dc), static_cast<const __nv_bfloat16*>(y),
                                  static_cast<const __nv_bfloat16*>(u), mean, rstd, bn_w, bn_b, sums, inv_count, wdw,
                                  static_cast<__nv_bfloat16*>(du), dwdw, dbdw, dbn_w, dbn_b, T, d, k);
  };
  if (bn_train) args(convmod_bwd_kernel<true>); else args(convmod_bwd_kernel<false>);
  B2C_CUDA(cudaGetLastError());
  return 0;
}
