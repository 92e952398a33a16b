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

// stage GLU(u) for frames [t_lo, t_lo + nrow) of utterance b into smem (zero outside [0,T))
__device__ __forceinline__ void stage_glu(const __nv_bfloat16* __restrict__ u, float2* __restrict__ sG, int b, int T,
                                          int d, int c0, int t_lo, int nrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = c0 + lane * 2;
  for (int r = warp; r < nrow; r += 8) {
    const int t = t_lo + r;
    float2 g = make_float2(0.f, 0.f);
    if (t >= 0 && t < T && ch < d) {
      const __nv_bfloat16* row = u + (static_cast<int64_t>(b) * T + t) * (2 * d);
      const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(row + ch));
      const float2 gt = unpack_bf16(*reinterpret_cast<const uint32_t*>(row + d + ch));
      g.x = a.x * sigmoid_f(gt.x);
      g.y = a.y * sigmoid_f(gt.y);
    }
    sG[r * 32 + lane] = g;
  }
}

// MODE 0: eval  -> out = SiLU(BN_running(y))            (bn_scale/bn_shift folded per channel)
// MODE 1: train -> y (bf16) + per-channel sum / sumsq   (stats fp64 [2][d], atomics once per CTA per channel)
template <int MODE>
__global__ void __launch_bounds__(256) convmod_fwd_kernel(const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ wdw, const float* __restrict__ bdw,
                                                          const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                          const float* __restrict__ bn_mean,
                                                          const float* __restrict__ bn_var, float eps,
                                                          __nv_bfloat16* __restrict__ out, double* __restrict__ stats,
                                                          int T, int d, int k) {
  extern __shared__ float2 sG[];                 // [(CM_TT + k - 1)][32] float2
  __shared__ float red[2][8][CM_CH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t0 = blockIdx.x * CM_TT, c0 = blockIdx.y * CM_CH, b = blockIdx.z;
  const int half = (k - 1) / 2;
  const int ch = c0 + lane * 2;
  const bool ch_ok = ch < d;

  stage_glu(u, sG, b, T, d, c0, t0 - half, CM_TT + k - 1);

  float w0[CM_MAXK], w1[CM_MAXK];
#pragma unroll
  for (int j = 0; j < CM_MAXK; ++j) {
    w0[j] = (j < k && ch_ok) ? __ldg(wdw + static_cast<int64_t>(ch) * k + j) : 0.f;
    w1[j] = (j < k && ch_ok) ? __ldg(wdw + static_cast<int64_t>(ch + 1) * k + j) : 0.f;
  }
  const float bias0 = ch_ok ? __ldg(bdw + ch) : 0.f, bias1 = ch_ok ? __ldg(bdw + ch + 1) : 0.f;
  float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
  if (MODE == 0 && ch_ok) {
    sc0 = __ldg(bn_w + ch) * rsqrtf(__ldg(bn_var + ch) + eps);
    sc1 = __ldg(bn_w + ch + 1) * rsqrtf(__ldg(bn_var + ch + 1) + eps);
    sh0 = __ldg(bn_b + ch) - __ldg(bn_mean + ch) * sc0;
    sh1 = __ldg(bn_b + ch + 1) - __ldg(bn_mean + ch + 1) * sc1;
  }
  __syncthreads();

  float s_sum0 = 0.f, s_sum1 = 0.f, s_sq0 = 0.f, s_sq1 = 0.f;
#pragma unroll 1
  for (int grp = 0; grp < (CM_TT / 8) / CM_G; ++grp) {          // 16 frames per warp = 2 groups of 8
    const int tl = warp * (CM_TT / 8) + grp * CM_G;             // first local output frame of the group
    float a0[CM_G], a1[CM_G];
#pragma unroll
    for (int i = 0; i < CM_G; ++i) { a0[i] = bias0; a1[i] = bias1; }
    // window rows tl .. tl + CM_G + k - 2 ; output i uses rows tl+i .. tl+i+k-1 with taps 0..k-1
#pragma unroll
    for (int r = 0; r < CM_G + CM_MAXK - 1; ++r) {
      if (r < CM_G + k - 1) {
        const float2 g = sG[(tl + r) * 32 + lane];
#pragma unroll
        for (int i = 0; i < CM_G; ++i) {
          const int j = r - i;
          if (j >= 0 && j < CM_MAXK) { a0[i] = fmaf(g.x, w0[j], a0[i]); a1[i] = fmaf(g.y, w1[j], a1[i]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CM_G; ++i) {
      const int t = t0 + tl + i;
      if (t < T && ch_ok) {
        float y0 = a0[i], y1 = a1[i];
        if (MODE == 1) {
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
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2);
  convmod_fwd_kernel<0><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, bn_w, bn_b, bn_mean, bn_var, eps,
      static_cast<__nv_bfloat16*>(out), nullptr, T, d, k);
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convmod_dw_fwd(const void* u, const float* wdw, const float* bdw, void* y, double* stats, int B,
                                  int T, int d, int k, void* stream) {
  B2C_REQUIRE(u && wdw && bdw && y && stats, "b2c_convmod_dw_fwd: null pointer");
  if (int rc = check_convmod(B, T, d, k, "b2c_convmod_dw_fwd")) return rc;
  dim3 grid((T + CM_TT - 1) / CM_TT, (d + CM_CH - 1) / CM_CH, B);
  const size_t smem = static_cast<size_t>(CM_TT + k - 1) * 32 * sizeof(float2);
  convmod_fwd_kernel<1><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(u), wdw, bdw, nullptr, nullptr, nullptr, nullptr, 0.f,
      static_cast<__nv_bfloat16*>(y), stats, T, d, k);
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
