// LayerNorm forward / backward and small memory-bound helpers (HBM/L2-bound, warp-shuffle reductions).
// Replaces torch.nn.LayerNorm at torchaudio/models/conformer.py:42,103,151,165 (5 per layer).
//
// One warp per row: a row of d <= 1024 fp32 lives in registers as float4 chunks (chunk index = lane + 32*i),
// so global accesses are 512-byte coalesced per warp instruction; statistics use the two-pass formula
// (mean, then centred sum of squares) entirely in registers -- x is read from HBM exactly once.
// Algorithmic bytes per row: fwd 4d (read) + 2d or 4d (write); bwd 4d (x) + 2d/4d (dy) + 4d (+4d when accumulating).
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int LN_MAX_CHUNKS = 8;  // float4 chunks per lane -> d <= 1024 (kernels are instantiated for 2, 4, 8)
constexpr int LN_WARPS = 8;

template <bool OUT_BF16, int LN_MAXV>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, void* __restrict__ y,
                                                                      float* __restrict__ mean_out,
                                                                      float* __restrict__ rstd_out, int64_t rows, int d,
                                                                      float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      v[i] = xr[c];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, e = v[i].w - mean;
      ss += a * a + b * b + cc * cc + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (OUT_BF16) {
        reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + row * d)[c] =
            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      } else {
        reinterpret_cast<float4*>(static_cast<float*>(y) + row * d)[c] = o;
      }
    }
  }
}

// dx = rstd * (g - mean_d(g) - xhat * mean_d(g*xhat)),  g = dy*gamma ; dgamma += dy*xhat ; dbeta += dy
// CAST: additionally emits  cast_out = bf16(alpha * dropout(dx_new))  and its column sums -- the operand (and bias
// gradient) the NEXT sub-layer's backward starts from (what b2c_scale_cast would compute in a second pass over dx).
struct LnCast {
  __nv_bfloat16* out;
  float* colsum;
  float alpha;
  DropCfg drop;
};
// d <= 256 (LN_MAXV = 2): capped at 80 registers so that THREE CTAs stay resident per SM (two at 128 registers for d <= 512),
// which is what the launch geometry below assumes.  Uncapped, the fp32-dy + cast instantiation (the final LayerNorm of a layer) took 83 -> 88 allocated
// registers -> two CTAs per SM -> its single wave of 444 CTAs ran as 1.5 waves: 28.3 us against 17.6 us for the bf16-dy
// instantiation on the same rows (tools/kernel_bench.py ln).
template <bool DY_BF16, int LN_MAXV, bool CAST>
__global__ void __launch_bounds__(LN_WARPS * 32, LN_MAXV <= 2 ? 3 : (LN_MAXV <= 4 ? 2 : 1)) layernorm_bwd_kernel(const void* __restrict__ dy,
                                                                      const float* __restrict__ x,
                                                                      const float* __restrict__ gamma,
                                                                      const float* __restrict__ mean_in,
                                                                      const float* __restrict__ rstd_in,
                                                                      float* __restrict__ dx, int accumulate,
                                                                      float* __restrict__ dgamma,
                                                                      float* __restrict__ dbeta, int64_t rows, int d,
                                                                      int rows_per_warp, LnCast cast) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];           // [2][d] per-block partial dgamma / dbeta (+ [d] cast column sums)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = d >> 2;
  for (int i = threadIdx.x; i < (CAST ? 3 : 2) * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 dg[LN_MAXV], db[LN_MAXV], cs[CAST ? LN_MAXV : 1];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (CAST ? LN_MAXV : 1); ++i) cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  DropCfg drop = cast.drop;
  if (CAST && drop.thresh16) drop = drop_resolve(drop);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp) * rows_per_warp;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int64_t row = row0 + rr;
    if (row >= rows) break;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
    float4 xh[LN_MAXV], g[LN_MAXV], prev[LN_MAXV];
    float s1 = 0.f, s2 = 0.f;
    float4* dxr = reinterpret_cast<float4*>(dx + row * d);
    // the running dx (residual-stream gradient) is fetched together with x and dy: one memory round trip per row
    // instead of two (it used to be read only after the two reductions)
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < LN_MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) prev[i] = dxr[c];
      }
    }
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 xv = xr[c];
        float4 dyv;
        if constexpr (DY_BF16) {
          const uint2 u = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy) + row * d)[c];
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
          dyv = make_float4(a.x, a.y, b.x, b.y);
        } else {
          dyv = reinterpret_cast<const float4*>(static_cast<const float*>(dy) + row * d)[c];
        }
        const float4 gm = __ldg(g4 + c);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(dyv.x * gm.x, dyv.y * gm.y, dyv.z * gm.z, dyv.w * gm.w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg[i].x += dyv.x * xh[i].x; dg[i].y += dyv.y * xh[i].y; dg[i].z += dyv.z * xh[i].z; dg[i].w += dyv.w * xh[i].w;
        db[i].x += dyv.x; db[i].y += dyv.y; db[i].z += dyv.z; db[i].w += dyv.w;
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 o;
        o.x = rstd * (g[i].x - s1 - xh[i].x * s2);
        o.y = rstd * (g[i].y - s1 - xh[i].y * s2);
        o.z = rstd * (g[i].z - s1 - xh[i].z * s2);
        o.w = rstd * (g[i].w - s1 - xh[i].w * s2);
        if (accumulate) {
          o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w;
        }
        dxr[c] = o;
        if constexpr (CAST) {
          float f[4] = {o.x * cast.alpha, o.y * cast.alpha, o.z * cast.alpha, o.w * cast.alpha};
          const uint64_t idx = static_cast<uint64_t>(row) * d + 4 * c;
          if (drop.thresh16) {
            float s0, s1, s2, s3;
            drop_scale4g(drop, idx, s0, s1, s2, s3);
            f[0] *= s0; f[1] *= s1; f[2] *= s2; f[3] *= s3;
          }
          const uint2 packed = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
          *reinterpret_cast<uint2*>(cast.out + idx) = packed;
          const float2 p0 = unpack_bf16(packed.x), p1 = unpack_bf16(packed.y);   // sums of the ROUNDED values
          cs[i].x += p0.x; cs[i].y += p0.y; cs[i].z += p1.x; cs[i].w += p1.y;
        }
      }
    }
  }
  // block-level reduction of the parameter gradients, then one atomic per channel per block
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      atomicAdd(&red[4 * c + 0], dg[i].x); atomicAdd(&red[4 * c + 1], dg[i].y);
      atomicAdd(&red[4 * c + 2], dg[i].z); atomicAdd(&red[4 * c + 3], dg[i].w);
      atomicAdd(&red[d + 4 * c + 0], db[i].x); atomicAdd(&red[d + 4 * c + 1], db[i].y);
      atomicAdd(&red[d + 4 * c + 2], db[i].z); atomicAdd(&red[d + 4 * c + 3], db[i].w);
      if constexpr (CAST) {
        atomicAdd(&red[2 * d + 4 * c + 0], cs[i].x); atomicAdd(&red[2 * d + 4 * c + 1], cs[i].y);
        atomicAdd(&red[2 * d + 4 * c + 2], cs[i].z); atomicAdd(&red[2 * d + 4 * c + 3], cs[i].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[d + i]);
    if constexpr (CAST) {
      if (cast.colsum) atomicAdd(cast.colsum + i, red[2 * d + i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm pairs
// Between two Conformer layers the stream passes through TWO LayerNorms back to back: the final LayerNorm of layer i
// (conformer.py:211, fp32 output = the next layer's residual stream) and the LayerNorm of layer i+1's first feed-forward
// module (conformer.py:103, bf16 output = a GEMM operand).  Run separately the second one re-reads what the first has just
// written; the pair kernels keep the row in registers: y1 = LN1(x) is stored AND normalised again, a2 = LN2(y1).
// Same instruction sequence per row as two layernorm_fwd_kernel launches, so the results are bit-identical.
template <int LN_MAXV>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_pair_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma1, const float* __restrict__ beta1, float* __restrict__ y1,
    float* __restrict__ mean1_out, float* __restrict__ rstd1_out, const float* __restrict__ gamma2,
    const float* __restrict__ beta2, __nv_bfloat16* __restrict__ a2, float* __restrict__ mean2_out,
    float* __restrict__ rstd2_out, int64_t rows, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      v[i] = xr[c];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, e = v[i].w - mean;
      ss += a * a + b * b + cc * cc + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / d + eps);
  if (lane == 0) {
    if (mean1_out) mean1_out[row] = mean;
    if (rstd1_out) rstd1_out[row] = rstd;
  }
  s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma1) + c), b = __ldg(reinterpret_cast<const float4*>(beta1) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      reinterpret_cast<float4*>(y1 + row * d)[c] = o;
      v[i] = o;                                        // the second LayerNorm's input
      s += o.x + o.y + o.z + o.w;
    }
  }
  const float mean2 = warp_sum(s) / d;
  ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float a = v[i].x - mean2, b = v[i].y - mean2, cc = v[i].z - mean2, e = v[i].w - mean2;
      ss += a * a + b * b + cc * cc + e * e;
    }
  }
  const float rstd2 = rsqrtf(warp_sum(ss) / d + eps);
  if (lane == 0) {
    if (mean2_out) mean2_out[row] = mean2;
    if (rstd2_out) rstd2_out[row] = rstd2;
  }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma2) + c), b = __ldg(reinterpret_cast<const float4*>(beta2) + c);
      float4 o;
      o.x = (v[i].x - mean2) * rstd2 * g.x + b.x;
      o.y = (v[i].y - mean2) * rstd2 * g.y + b.y;
      o.z = (v[i].z - mean2) * rstd2 * g.z + b.z;
      o.w = (v[i].w - mean2) * rstd2 * g.w + b.w;
      reinterpret_cast<uint2*>(a2 + row * d)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

// Backward of the pair, one pass: the gradient arriving at a2 (bf16, from the feed-forward module's dgrad GEMM) goes
// through LN2 and is added to the running residual-stream gradient `prev` (what layernorm_bwd_kernel<true,..,false> does
// with accumulate = 1); the sum goes through LN1 and leaves as dx (fp32) and as cast_out = bf16(alpha * dropout(dx)) with
// its column sums (what layernorm_bwd_kernel<false,..,true> does): the intermediate gradient is neither written nor re-read.
struct LnPairBwd {
  const __nv_bfloat16* da2;                            // gradient w.r.t. a2
  const float* y1;   const float* gamma2; const float* mean2; const float* rstd2; float* dgamma2; float* dbeta2;
  const float* prev;                                   // running gradient w.r.t. y1 (read only)
  const float* x;    const float* gamma1; const float* mean1; const float* rstd1; float* dgamma1; float* dbeta1;
  float* dx;                                           // out: gradient w.r.t. x
};
template <int LN_MAXV>
__global__ void __launch_bounds__(LN_WARPS * 32, LN_MAXV <= 2 ? 2 : 1) layernorm_bwd_pair_kernel(const LnPairBwd a, int64_t rows, int d,
                                                                                                 int rows_per_warp, LnCast cast) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];           // [5][d]: dgamma2, dbeta2, dgamma1, dbeta1, cast column sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = d >> 2;
  for (int i = threadIdx.x; i < 5 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 dg2[LN_MAXV], db2[LN_MAXV], dg1[LN_MAXV], db1[LN_MAXV], cs[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) dg2[i] = db2[i] = dg1[i] = db1[i] = cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  DropCfg drop = cast.drop;
  if (drop.thresh16) drop = drop_resolve(drop);
  const float4* g2p = reinterpret_cast<const float4*>(a.gamma2);
  const float4* g1p = reinterpret_cast<const float4*>(a.gamma1);
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp) * rows_per_warp;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int64_t row = row0 + rr;
    if (row >= rows) break;
    const float mean2 = a.mean2[row], rstd2 = a.rstd2[row], mean1 = a.mean1[row], rstd1 = a.rstd1[row];
    float4 xh[LN_MAXV], g[LN_MAXV], t[LN_MAXV], xf[LN_MAXV];
    // every operand of the row is fetched up front: one memory round trip for both LayerNorms
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        t[i] = reinterpret_cast<const float4*>(a.prev + row * d)[c];
        xf[i] = reinterpret_cast<const float4*>(a.x + row * d)[c];
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 yv = reinterpret_cast<const float4*>(a.y1 + row * d)[c];
        const uint2 u = reinterpret_cast<const uint2*>(a.da2 + row * d)[c];
        const float2 p = unpack_bf16(u.x), q = unpack_bf16(u.y);
        const float4 dyv = make_float4(p.x, p.y, q.x, q.y);
        const float4 gm = __ldg(g2p + c);
        xh[i] = make_float4((yv.x - mean2) * rstd2, (yv.y - mean2) * rstd2, (yv.z - mean2) * rstd2, (yv.w - mean2) * rstd2);
        g[i] = make_float4(dyv.x * gm.x, dyv.y * gm.y, dyv.z * gm.z, dyv.w * gm.w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg2[i].x += dyv.x * xh[i].x; dg2[i].y += dyv.y * xh[i].y; dg2[i].z += dyv.z * xh[i].z; dg2[i].w += dyv.w * xh[i].w;
        db2[i].x += dyv.x; db2[i].y += dyv.y; db2[i].z += dyv.z; db2[i].w += dyv.w;
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
    float u1 = 0.f, u2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        // gradient w.r.t. y1 = LN2 backward + the running gradient (exactly the accumulate = 1 path)
        float4 o;
        o.x = rstd2 * (g[i].x - s1 - xh[i].x * s2);
        o.y = rstd2 * (g[i].y - s1 - xh[i].y * s2);
        o.z = rstd2 * (g[i].z - s1 - xh[i].z * s2);
        o.w = rstd2 * (g[i].w - s1 - xh[i].w * s2);
        o.x += t[i].x; o.y += t[i].y; o.z += t[i].z; o.w += t[i].w;
        // ... which is the dy of LN1
        const float4 gm = __ldg(g1p + c);
        xh[i] = make_float4((xf[i].x - mean1) * rstd1, (xf[i].y - mean1) * rstd1, (xf[i].z - mean1) * rstd1, (xf[i].w - mean1) * rstd1);
        g[i] = make_float4(o.x * gm.x, o.y * gm.y, o.z * gm.z, o.w * gm.w);
        u1 += g[i].x + g[i].y + g[i].z + g[i].w;
        u2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg1[i].x += o.x * xh[i].x; dg1[i].y += o.y * xh[i].y; dg1[i].z += o.z * xh[i].z; dg1[i].w += o.w * xh[i].w;
        db1[i].x += o.x; db1[i].y += o.y; db1[i].z += o.z; db1[i].w += o.w;
      }
    }
    u1 = warp_sum(u1) / d;
    u2 = warp_sum(u2) / d;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 o;
        o.x = rstd1 * (g[i].x - u1 - xh[i].x * u2);
        o.y = rstd1 * (g[i].y - u1 - xh[i].y * u2);
        o.z = rstd1 * (g[i].z - u1 - xh[i].z * u2);
        o.w = rstd1 * (g[i].w - u1 - xh[i].w * u2);
        reinterpret_cast<float4*>(a.dx + row * d)[c] = o;
        float f[4] = {o.x * cast.alpha, o.y * cast.alpha, o.z * cast.alpha, o.w * cast.alpha};
        const uint64_t idx = static_cast<uint64_t>(row) * d + 4 * c;
        if (drop.thresh16) {
          float k0, k1, k2, k3;
          drop_scale4g(drop, idx, k0, k1, k2, k3);
          f[0] *= k0; f[1] *= k1; f[2] *= k2; f[3] *= k3;
        }
        const uint2 packed = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
        *reinterpret_cast<uint2*>(cast.out + idx) = packed;
        const float2 p0 = unpack_bf16(packed.x), p1 = unpack_bf16(packed.y);   // sums of the ROUNDED values
        cs[i].x += p0.x; cs[i].y += p0.y; cs[i].z += p1.x; cs[i].w += p1.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 acc[5] = {dg2[i], db2[i], dg1[i], db1[i], cs[i]};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        atomicAdd(&red[k * d + 4 * c + 0], acc[k].x); atomicAdd(&red[k * d + 4 * c + 1], acc[k].y);
        atomicAdd(&red[k * d + 4 * c + 2], acc[k].z); atomicAdd(&red[k * d + 4 * c + 3], acc[k].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(a.dgamma2 + i, red[i]);
    atomicAdd(a.dbeta2 + i, red[d + i]);
    atomicAdd(a.dgamma1 + i, red[2 * d + i]);
    atomicAdd(a.dbeta1 + i, red[3 * d + i]);
    if (cast.colsum) atomicAdd(cast.colsum + i, red[4 * d + i]);
  }
}

// ---------------------------------------------------------------------------------------------- helpers
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i);
    const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
    *reinterpret_cast<uint4*>(dst + i) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// table entries: {src fp32*, dst bf16*, n}; blockIdx.y = tensor, grid-stride over 8-element chunks
struct CastEntry { const float* src; __nv_bfloat16* dst; int64_t n; };
__global__ void multi_cast_kernel(const CastEntry* __restrict__ table) {
  pdl_trigger();
  pdl_wait();
  const CastEntry e = table[blockIdx.y];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < e.n; i += stride) {
    if (i + 8 <= e.n) {
      const float4 a = *reinterpret_cast<const float4*>(e.src + i);
      const float4 b = *reinterpret_cast<const float4*>(e.src + i + 4);
      *reinterpret_cast<uint4*>(e.dst + i) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    } else {
      for (int64_t j = i; j < e.n; ++j) e.dst[j] = __float2bfloat16(e.src[j]);
    }
  }
}

// dst = bf16(alpha * dropout(src)); optionally colsum[c] += sum_r dst[r, c].  2-D mapping: a thread owns 4 adjacent
// columns and walks rows (coalesced 16 B loads across the row), so column partials live in registers; one shared-memory
// reduction and one atomic per column per CTA.
__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        int64_t rows, int cols, float alpha, DropCfg drop,
                                                        float* __restrict__ colsum, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float cs_red[];                    // [cols] column partials of this CTA
  const int c4 = cols >> 2;                            // float4 groups per row
  const int tpr = c4 < 256 ? c4 : 256;                 // threads per row
  const int rstep = 256 / tpr;                         // rows per pass
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  if (colsum) {
    for (int i = threadIdx.x; i < cols; i += 256) cs_red[i] = 0.f;
    __syncthreads();
  }
  if (drop.thresh16) drop = drop_resolve(drop);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  if (tr < rstep) {
    for (int cg = tc; cg < c4; cg += tpr) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int64_t r = r0 + tr; r < r1; r += rstep) {
        const int64_t i = r * cols + cg * 4;
        const float4 a = *reinterpret_cast<const float4*>(src + i);
        float f[4] = {a.x * alpha, a.y * alpha, a.z * alpha, a.w * alpha};
        if (drop.thresh16) {
          float s0, s1, s2, s3;
          drop_scale4g(drop, static_cast<uint64_t>(i), s0, s1, s2, s3);
          f[0] *= s0; f[1] *= s1; f[2] *= s2; f[3] *= s3;
        }
        const uint2 packed = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
        *reinterpret_cast<uint2*>(dst + i) = packed;
        if (colsum) {                                  // sums of the ROUNDED values: exactly what the wgrad GEMM will see
          const float2 p0 = unpack_bf16(packed.x), p1 = unpack_bf16(packed.y);
          a0 += p0.x; a1 += p0.y; a2 += p1.x; a3 += p1.y;
        }
      }
      if (colsum) {
        atomicAdd(&cs_red[cg * 4 + 0], a0); atomicAdd(&cs_red[cg * 4 + 1], a1);
        atomicAdd(&cs_red[cg * 4 + 2], a2); atomicAdd(&cs_red[cg * 4 + 3], a3);
      }
    }
  }
  if (colsum) {
    __syncthreads();
    for (int i = threadIdx.x; i < cols; i += 256) atomicAdd(colsum + i, cs_red[i]);
  }
}

// column sums of a bf16 matrix: block = 32 x 8 threads, each thread owns 2 adjacent columns over a row slab
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, int64_t ld, int64_t rows,
                                                         int64_t cols, float* __restrict__ out, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  // thread = one fixed group of 8 columns walking down this CTA's rows with 16-byte loads, four rows in flight; the
  // CTA's partials meet in shared memory and leave as one atomic per column (the first version read 4 bytes per thread
  // per step: 20 us for 37 MB)
  extern __shared__ float cs_dyn[];                      // [cols]
  const int cg = static_cast<int>(cols >> 3);
  const int tpr = cg < 256 ? cg : 256, rpi = 256 / tpr;
  const int tr = threadIdx.x / tpr, tg = threadIdx.x - tr * tpr;
  for (int i = threadIdx.x; i < cols; i += 256) cs_dyn[i] = 0.f;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  if (tr < rpi) {
    for (int g = tg; g < cg; g += tpr) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
      constexpr int UN = 8;
      for (int64_t r = r0 + tr; r < r1; r += static_cast<int64_t>(UN) * rpi) {
        uint4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int64_t rr = r + static_cast<int64_t>(u) * rpi;
          v[u] = rr < r1 ? *reinterpret_cast<const uint4*>(a + rr * ld + g * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const float2 p0 = unpack_bf16(v[u].x), p1 = unpack_bf16(v[u].y), p2 = unpack_bf16(v[u].z), p3 = unpack_bf16(v[u].w);
          s[0] += p0.x; s[1] += p0.y; s[2] += p1.x; s[3] += p1.y; s[4] += p2.x; s[5] += p2.y; s[6] += p3.x; s[7] += p3.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&cs_dyn[g * 8 + j], s[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cols; i += 256) atomicAdd(out + i, cs_dyn[i]);
}

// generic fallback (any even cols / ld): a lane owns two adjacent columns
__global__ void __launch_bounds__(256) colsum_bf16_pair_kernel(const __nv_bfloat16* __restrict__ a, int64_t ld, int64_t rows,
                                                              int64_t cols, float* __restrict__ out, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = static_cast<int64_t>(blockIdx.x) * 64 + tx * 2;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (col < cols) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(a + r * ld + col));
      s0 += v.x; s1 += v.y;
    }
  }
  red[ty][tx * 2] = s0;
  red[ty][tx * 2 + 1] = s1;
  __syncthreads();
  if (ty == 0) {
    for (int j = 1; j < 8; ++j) { s0 += red[j][tx * 2]; s1 += red[j][tx * 2 + 1]; }
    if (col < cols) { atomicAdd(out + col, s0); atomicAdd(out + col + 1, s1); }
  }
}

__global__ void repack_heads_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                    int nblk, int src_dh, int dst_dh) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_row = static_cast<int64_t>(nblk) * dst_dh;
  if (i >= rows * per_row) return;
  const int64_t r = i / per_row;
  const int rem = static_cast<int>(i % per_row);
  const int h = rem / dst_dh, c = rem % dst_dh;
  dst[i] = c < src_dh ? src[r * nblk * src_dh + h * src_dh + c] : __float2bfloat16(0.f);
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                                 float* mean, float* rstd, int64_t rows, int d, float eps, void* stream) {
  B2C_REQUIRE(x && gamma && beta && y, "b2c_layernorm_fwd: null pointer");
  B2C_REQUIRE(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_CHUNKS, "b2c_layernorm_fwd: d=%d must be a multiple of 4, <= 1024", d);
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (d / 4 + 31) / 32;
  const unsigned grid = static_cast<unsigned>((rows + LN_WARPS - 1) / LN_WARPS);
#define B2C_LN_FWD(BF, NV) B2C_CUDA(launch_k(layernorm_fwd_kernel<BF, NV>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, y, mean, rstd, rows, d, eps))
  if (y_dtype == B2C_BF16) { if (nv <= 2) B2C_LN_FWD(true, 2); else if (nv <= 4) B2C_LN_FWD(true, 4); else B2C_LN_FWD(true, 8); }
  else                     { if (nv <= 2) B2C_LN_FWD(false, 2); else if (nv <= 4) B2C_LN_FWD(false, 4); else B2C_LN_FWD(false, 8); }
#undef B2C_LN_FWD
  B2C_CUDA(cudaGetLastError());
  return 0;
}

static int layernorm_bwd_launch(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean,
                                const float* rstd, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                                int64_t rows, int d, const LnCast* cast, void* stream) {
  B2C_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "b2c_layernorm_bwd: null pointer");
  B2C_REQUIRE(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_CHUNKS, "b2c_layernorm_bwd: d=%d must be a multiple of 4, <= 1024", d);
  if (rows <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (d / 4 + 31) / 32;
  LnCast c0 = {};
  // ONE wave of CTAs (3 / 2 / 1 co-resident per SM for d <= 256 / 512 / 1024: the kernel's launch bounds cap its registers at
  // 80 / 128 / 255): each warp walks a contiguous run of rows, so the parameter-gradient atomics stay rare and no partial
  // last wave trails behind (995 CTAs on 444 slots used to run as 2.24 waves)
  const int per_sm = nv <= 2 ? 3 : (nv <= 4 ? 2 : 1);
  const int64_t target_warps = static_cast<int64_t>(num_sms()) * per_sm * LN_WARPS;
  int rows_per_warp = static_cast<int>((rows + target_warps - 1) / target_warps);
  if (rows_per_warp < 1) rows_per_warp = 1;
  const int64_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const unsigned grid = static_cast<unsigned>((warps + LN_WARPS - 1) / LN_WARPS);
  const size_t smem = (cast ? 3 : 2) * d * sizeof(float);
#define B2C_LN_BWD(BF, NV)                                                                                      \
  do {                                                                                                          \
    if (cast) B2C_CUDA(launch_k(layernorm_bwd_kernel<BF, NV, true>, dim3(grid), dim3(LN_WARPS * 32), smem, st,                            \
        dy, x, gamma, mean, rstd, dx, dx_accumulate, dgamma, dbeta, rows, d, rows_per_warp, *cast));            \
    else B2C_CUDA(launch_k(layernorm_bwd_kernel<BF, NV, false>, dim3(grid), dim3(LN_WARPS * 32), smem, st,                                \
        dy, x, gamma, mean, rstd, dx, dx_accumulate, dgamma, dbeta, rows, d, rows_per_warp, c0));               \
  } while (0)
  if (dy_dtype == B2C_BF16) { if (nv <= 2) B2C_LN_BWD(true, 2); else if (nv <= 4) B2C_LN_BWD(true, 4); else B2C_LN_BWD(true, 8); }
  else                      { if (nv <= 2) B2C_LN_BWD(false, 2); else if (nv <= 4) B2C_LN_BWD(false, 4); else B2C_LN_BWD(false, 8); }
#undef B2C_LN_BWD
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                                 int64_t rows, int d, void* stream) {
  return layernorm_bwd_launch(dy, dy_dtype, x, gamma, mean, rstd, dx, dx_accumulate, dgamma, dbeta, rows, d, nullptr, stream);
}

extern "C" int b2c_layernorm_bwd_cast(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean,
                                      const float* rstd, float* dx, int dx_accumulate, float* dgamma, float* dbeta,
                                      int64_t rows, int d, void* cast_out, float alpha, float drop_p, uint64_t drop_seed,
                                      float* colsum, void* stream) {
  B2C_REQUIRE(cast_out, "b2c_layernorm_bwd_cast: null cast_out");
  LnCast c;
  c.out = static_cast<__nv_bfloat16*>(cast_out); c.colsum = colsum; c.alpha = alpha;
  c.drop = make_drop(drop_p, drop_seed);
  return layernorm_bwd_launch(dy, dy_dtype, x, gamma, mean, rstd, dx, dx_accumulate, dgamma, dbeta, rows, d, &c, stream);
}

namespace b2c {
// out[i] = x[i] + e[i]  (fp32, float4): `x = x + rel_pos_emb` in front of the attention LayerNorm of the reference's own
// positional variant (/root/reference/unused_lib/conformer.py:101-103; table = RelativePositionalEmbeddings, one row per
// flattened token).  n elements; e holds at least n.
__global__ void __launch_bounds__(256) add_f32_kernel(const float4* __restrict__ x, const float4* __restrict__ e,
                                                     float4* __restrict__ out, int64_t n4) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 a = x[i], b = e[i];
    out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

}  // namespace b2c
extern "C" int b2c_add_f32(const float* x, const float* e, float* out, int64_t n, void* stream) {
  B2C_REQUIRE(x && e && out && n > 0 && n % 4 == 0, "b2c_add_f32: null pointer or n not a multiple of 4");
  B2C_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "b2c_add_f32: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  const int grid = static_cast<int>(n4 / 256 + 1 < 8 * 148 ? n4 / 256 + 1 : 8 * 148);
  B2C_CUDA(launch_k(b2c::add_f32_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float4*>(x),
                                                                         reinterpret_cast<const float4*>(e),
                                                                         reinterpret_cast<float4*>(out), n4));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int b2c_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  B2C_REQUIRE(src && dst, "b2c_cast_f32_bf16: null pointer");
  if (n <= 0) return 0;
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "b2c_cast_f32_bf16: pointers must be 16-byte aligned");
  const int64_t threads = (n + 7) / 8;
  B2C_CUDA(launch_k(cast_f32_bf16_kernel, dim3(static_cast<unsigned>((threads + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      src, static_cast<__nv_bfloat16*>(dst), n));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_multi_cast_f32_bf16(const void* table_dev, int n_tensors, int64_t max_n, void* stream) {
  B2C_REQUIRE(table_dev && n_tensors > 0, "b2c_multi_cast_f32_bf16: bad argument");
  int64_t gx = (max_n / 8 + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  B2C_CUDA(launch_k(multi_cast_kernel, dim3(dim3(static_cast<unsigned>(gx), n_tensors)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const CastEntry*>(table_dev)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_scale_cast(const float* src, void* dst, int64_t rows, int64_t cols, float alpha, float drop_p,
                              uint64_t drop_seed, float* colsum, void* stream) {
  B2C_REQUIRE(src && dst, "b2c_scale_cast: null pointer");
  if (rows <= 0 || cols <= 0) return 0;
  B2C_REQUIRE(cols % 4 == 0 && cols <= 8192, "b2c_scale_cast: cols must be a multiple of 4 (<= 8192)");
  const int64_t want_blocks = static_cast<int64_t>(num_sms()) * 4;
  int rows_per_block = static_cast<int>((rows + want_blocks - 1) / want_blocks);
  if (rows_per_block < 4) rows_per_block = 4;
  const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  B2C_CUDA(launch_k(scale_cast_kernel, dim3(grid), dim3(256), colsum ? cols * sizeof(float) : 0, static_cast<cudaStream_t>(stream), 
      src, static_cast<__nv_bfloat16*>(dst), rows, static_cast<int>(cols), alpha, make_drop(drop_p, drop_seed), colsum,
      rows_per_block));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_colsum_bf16(const void* a, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream) {
  B2C_REQUIRE(a && out, "b2c_colsum_bf16: null pointer");
  B2C_REQUIRE(cols % 2 == 0 && ld % 2 == 0, "b2c_colsum_bf16: cols and ld must be even");
  if (rows <= 0 || cols <= 0) return 0;
  if (cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && cols <= 8192) {
    const int cg = static_cast<int>(cols / 8), tpr = cg < 256 ? cg : 256, rpi = 256 / tpr;
    int64_t nblk = static_cast<int64_t>(num_sms()) * 4;
    int64_t rpb = (rows + nblk - 1) / nblk;
    rpb = ((rpb + rpi - 1) / rpi) * rpi;
    nblk = (rows + rpb - 1) / rpb;
    B2C_CUDA(launch_k(colsum_bf16_kernel, dim3(static_cast<unsigned>(nblk)), dim3(256), cols * sizeof(float), static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(a), ld, rows, cols, out, static_cast<int>(rpb)));
    B2C_CUDA(cudaGetLastError());
    return 0;
  }
  const unsigned gx = static_cast<unsigned>((cols + 63) / 64);
  int64_t want_y = (static_cast<int64_t>(num_sms()) * 4 + gx - 1) / gx;
  int rows_per_block = static_cast<int>((rows + want_y - 1) / want_y);
  rows_per_block = ((rows_per_block + 7) / 8) * 8;
  const unsigned gy = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
  B2C_CUDA(launch_k(colsum_bf16_pair_kernel, dim3(dim3(gx, gy)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(a), ld, rows, cols, out, rows_per_block));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_repack_heads(const void* src, void* dst, int64_t rows, int nblk, int src_dh, int dst_dh,
                                void* stream) {
  B2C_REQUIRE(src && dst && nblk > 0 && src_dh > 0 && dst_dh > 0, "b2c_repack_heads: bad argument");
  const int64_t n = rows * nblk * dst_dh;
  if (n <= 0) return 0;
  B2C_CUDA(launch_k(repack_heads_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, nblk, src_dh, dst_dh));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_layernorm_fwd_pair(const float* x, const float* gamma1, const float* beta1, float* y1, float* mean1,
                                      float* rstd1, const float* gamma2, const float* beta2, void* a2, float* mean2,
                                      float* rstd2, int64_t rows, int d, float eps, void* stream) {
  B2C_REQUIRE(x && gamma1 && beta1 && y1 && gamma2 && beta2 && a2, "b2c_layernorm_fwd_pair: null pointer");
  B2C_REQUIRE(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_CHUNKS, "b2c_layernorm_fwd_pair: d=%d must be a multiple of 4, <= 1024", d);
  if (rows <= 0) return 0;
  const unsigned grid = static_cast<unsigned>((rows + LN_WARPS - 1) / LN_WARPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2C_LN_PAIR(NV) B2C_CUDA(launch_k(layernorm_fwd_pair_kernel<NV>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, gamma1, beta1, y1, \
                                          mean1, rstd1, gamma2, beta2, static_cast<__nv_bfloat16*>(a2), mean2, rstd2, rows, d, eps))
  const int nv = (d / 4 + 31) / 32;
  if (nv <= 2) B2C_LN_PAIR(2); else if (nv <= 4) B2C_LN_PAIR(4); else B2C_LN_PAIR(8);
#undef B2C_LN_PAIR
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_layernorm_bwd_pair(const void* da2, const float* y1, const float* gamma2, const float* mean2,
                                      const float* rstd2, float* dgamma2, float* dbeta2, const float* prev, const float* x,
                                      const float* gamma1, const float* mean1, const float* rstd1, float* dgamma1,
                                      float* dbeta1, float* dx, int64_t rows, int d, void* cast_out, float alpha,
                                      float drop_p, uint64_t drop_seed, float* colsum, void* stream) {
  B2C_REQUIRE(da2 && y1 && gamma2 && mean2 && rstd2 && dgamma2 && dbeta2 && prev && x && gamma1 && mean1 && rstd1 && dgamma1 &&
                  dbeta1 && dx && cast_out, "b2c_layernorm_bwd_pair: null pointer");
  B2C_REQUIRE(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_CHUNKS, "b2c_layernorm_bwd_pair: d=%d must be a multiple of 4, <= 1024", d);
  if (rows <= 0) return 0;
  LnPairBwd a;
  a.da2 = static_cast<const __nv_bfloat16*>(da2);
  a.y1 = y1; a.gamma2 = gamma2; a.mean2 = mean2; a.rstd2 = rstd2; a.dgamma2 = dgamma2; a.dbeta2 = dbeta2;
  a.prev = prev;
  a.x = x; a.gamma1 = gamma1; a.mean1 = mean1; a.rstd1 = rstd1; a.dgamma1 = dgamma1; a.dbeta1 = dbeta1;
  a.dx = dx;
  LnCast c;
  c.out = static_cast<__nv_bfloat16*>(cast_out); c.colsum = colsum; c.alpha = alpha;
  c.drop = make_drop(drop_p, drop_seed);
  const int nv = (d / 4 + 31) / 32;
  const int per_sm = nv <= 2 ? 2 : 1;                 // launch bounds of the kernel: 128 registers for d <= 256
  const int64_t target_warps = static_cast<int64_t>(num_sms()) * per_sm * LN_WARPS;
  int rows_per_warp = static_cast<int>((rows + target_warps - 1) / target_warps);
  if (rows_per_warp < 1) rows_per_warp = 1;
  const int64_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const unsigned grid = static_cast<unsigned>((warps + LN_WARPS - 1) / LN_WARPS);
  const size_t smem = 5 * static_cast<size_t>(d) * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2C_LN_PAIRB(NV) B2C_CUDA(launch_k(layernorm_bwd_pair_kernel<NV>, dim3(grid), dim3(LN_WARPS * 32), smem, st, a, rows, d, rows_per_warp, c))
  if (nv <= 2) B2C_LN_PAIRB(2); else if (nv <= 4) B2C_LN_PAIRB(4); else B2C_LN_PAIRB(8);
#undef B2C_LN_PAIRB
  B2C_CUDA(cudaGetLastError());
  return 0;
}
