// BatchNorm1d over token rows + the SiLU that precedes it: the tail of the reference's projection block
//   x = projection_fc(x); x = swish(x); x = projection_batch_norm(x)        (/root/reference/lib/standard/asrnn.py:73-89,217)
// on [rows = batch_size * max_len, C = projection_out_size] (padded rows included, exactly as the reference's flatten(0,1)
// at asrnn.py:216 hands them to nn.BatchNorm1d).  The Linear + SiLU run as the tcgen05 GEMM with its SiLU epilogue
// (b2c_gemm_bf16, act = SILU, out2 = saved pre-activation); these kernels are the memory-bound rest:
//   bn1d_stats       per-channel sum / sum of squares of h (bf16) in fp64            (train mode)
//   (finalize = b2c_convmod_bn_finalize: mean, rstd, running statistics with momentum and unbiased variance)
//   bn1d_apply       y = (h - mean) * rstd * gamma + beta                            fp32 out
//   bn1d_bwd_reduce  S1 = sum dy, S2 = sum dy * hhat  (= dbeta, dgamma)
//   bn1d_bwd_apply   dh = gamma * rstd * (dy - S1/N - hhat * S2/N)  [eval: gamma * rstd * dy], then through the SiLU:
//                    dz = dh * SiLU'(z) (z = saved pre-activation), bf16 out + column sums of dz (the Linear's bias gradient)
// Thread = fixed group of 8 channels walking down rows (16-byte loads), per-CTA partial sums in shared memory, one fp64
// atomic per channel per CTA -- the scheme of the conv module's BatchNorm kernels (convmod.cu).
#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

__global__ void __launch_bounds__(256) bn1d_stats_kernel(const __nv_bfloat16* __restrict__ h, double* __restrict__ stats,
                                                        int64_t rows, int C, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];                       // [2][C]
  const int cg = C >> 3, rpi = 256 / cg;
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  __syncthreads();
  if (tr < rpi) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r < r_end; r += rpi) {
      const uint4 v = *reinterpret_cast<const uint4*>(h + r * C + tg * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(w[j]);
        s[2 * j] += f.x; s[2 * j + 1] += f.y;
        q[2 * j] = fmaf(f.x, f.x, q[2 * j]); q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[tg * 8 + j], s[j]);
      atomicAdd(&red[C + tg * 8 + j], q[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(stats + i, static_cast<double>(red[i]));
}

__global__ void __launch_bounds__(256) bn1d_apply_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y, int64_t rows,
                                                        int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t n8 = rows * (C >> 3);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (C >> 3)) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(h + i * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16(w[j]);
      const int c0 = c + 2 * j;
      const float a0 = __ldg(rstd + c0) * __ldg(gamma + c0), a1 = __ldg(rstd + c0 + 1) * __ldg(gamma + c0 + 1);
      o[2 * j] = fmaf(f.x - __ldg(mean + c0), a0, __ldg(beta + c0));
      o[2 * j + 1] = fmaf(f.y - __ldg(mean + c0 + 1), a1, __ldg(beta + c0 + 1));
    }
    *reinterpret_cast<float4*>(y + i * 8) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(y + i * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

__global__ void __launch_bounds__(256) bn1d_bwd_reduce_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             double* __restrict__ sums, int64_t rows, int C,
                                                             int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];                       // [2][C]
  const int cg = C >> 3, rpi = 256 / cg;
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  for (int i = threadIdx.x; i < 2 * C; i += 256) red[i] = 0.f;
  __syncthreads();
  if (tr < rpi) {
    float s1[8], s2[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] = s2[j] = 0.f;
      mu[j] = __ldg(mean + tg * 8 + j);
      rs[j] = __ldg(rstd + tg * 8 + j);
    }
    const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r < r_end; r += rpi) {
      const uint4 v = *reinterpret_cast<const uint4*>(h + r * C + tg * 8);
      const float4 d0 = *reinterpret_cast<const float4*>(dy + r * C + tg * 8);
      const float4 d1 = *reinterpret_cast<const float4*>(dy + r * C + tg * 8 + 4);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(w[j]);
        s1[2 * j] += d[2 * j]; s1[2 * j + 1] += d[2 * j + 1];
        s2[2 * j] = fmaf(d[2 * j], (f.x - mu[2 * j]) * rs[2 * j], s2[2 * j]);
        s2[2 * j + 1] = fmaf(d[2 * j + 1], (f.y - mu[2 * j + 1]) * rs[2 * j + 1], s2[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[tg * 8 + j], s1[j]);
      atomicAdd(&red[C + tg * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(sums + i, static_cast<double>(red[i]));
}

__global__ void __launch_bounds__(256) bn1d_bwd_apply_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                                                            const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const double* __restrict__ sums, double inv_count, int train,
                                                            __nv_bfloat16* __restrict__ dz, float* __restrict__ colsum,
                                                            int64_t rows, int C, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];                       // [C]
  const int cg = C >> 3, rpi = 256 / cg;
  const int tr = threadIdx.x / cg, tg = threadIdx.x - tr * cg;
  for (int i = threadIdx.x; i < C; i += 256) red[i] = 0.f;
  __syncthreads();
  if (tr < rpi) {
    float mu[8], rs[8], ga[8], a1[8], a2[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = tg * 8 + j;
      mu[j] = __ldg(mean + c); rs[j] = __ldg(rstd + c); ga[j] = __ldg(gamma + c) * rs[j];
      a1[j] = train ? static_cast<float>(sums[c] * inv_count) : 0.f;
      a2[j] = train ? static_cast<float>(sums[C + c] * inv_count) : 0.f;
      cs[j] = 0.f;
    }
    const int64_t r_end = min(rows, (static_cast<int64_t>(blockIdx.x) + 1) * rows_per_block);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_block + tr; r < r_end; r += rpi) {
      const uint4 vh = *reinterpret_cast<const uint4*>(h + r * C + tg * 8);
      const uint4 vz = *reinterpret_cast<const uint4*>(z + r * C + tg * 8);
      const float4 d0 = *reinterpret_cast<const float4*>(dy + r * C + tg * 8);
      const float4 d1 = *reinterpret_cast<const float4*>(dy + r * C + tg * 8 + 4);
      const uint32_t wh[4] = {vh.x, vh.y, vh.z, vh.w}, wz[4] = {vz.x, vz.y, vz.z, vz.w};
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fh = unpack_bf16(wh[j]), fz = unpack_bf16(wz[j]);
        const float hh0 = (fh.x - mu[2 * j]) * rs[2 * j], hh1 = (fh.y - mu[2 * j + 1]) * rs[2 * j + 1];
        const float dh0 = ga[2 * j] * (d[2 * j] - a1[2 * j] - hh0 * a2[2 * j]);
        const float dh1 = ga[2 * j + 1] * (d[2 * j + 1] - a1[2 * j + 1] - hh1 * a2[2 * j + 1]);
        o[j] = pack_bf16(dh0 * silu_grad_f(fz.x), dh1 * silu_grad_f(fz.y));
        const float2 rr = unpack_bf16(o[j]);                      // sums of the ROUNDED values: what the weight-gradient GEMM sees
        cs[2 * j] += rr.x; cs[2 * j + 1] += rr.y;
      }
      *reinterpret_cast<uint4*>(dz + r * C + tg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (colsum) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[tg * 8 + j], cs[j]);
    }
  }
  __syncthreads();
  if (colsum)
    for (int i = threadIdx.x; i < C; i += 256)
      if (red[i] != 0.f) atomicAdd(colsum + i, red[i]);
}

static int bn1d_grid(int64_t rows, int C, int64_t* rpb_out) {
  const int rpi = 256 / (C / 8);
  int64_t nblk = static_cast<int64_t>(num_sms()) * 2;
  int64_t rpb = (rows + nblk - 1) / nblk;
  rpb = ((rpb + rpi - 1) / rpi) * rpi;
  nblk = (rows + rpb - 1) / rpb;
  *rpb_out = rpb;
  return static_cast<int>(nblk);
}

}  // namespace b2c

using namespace b2c;

#define B2C_BN1D_CHECK(who)                                                                               \
  B2C_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, who ": C must be a multiple of 8, <= 2048, rows > 0")

extern "C" int b2c_bn1d_stats(const void* h, double* stats, int64_t rows, int C, void* stream) {
  B2C_REQUIRE(h && stats, "b2c_bn1d_stats: null pointer");
  B2C_BN1D_CHECK("b2c_bn1d_stats");
  int64_t rpb;
  const int grid = bn1d_grid(rows, C, &rpb);
  B2C_CUDA(launch_k(bn1d_stats_kernel, dim3(grid), dim3(256), 2 * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(h), stats, rows, C, static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_bn1d_apply(const void* h, const float* mean, const float* rstd, const float* gamma, const float* beta,
                              float* y, int64_t rows, int C, void* stream) {
  B2C_REQUIRE(h && mean && rstd && gamma && beta && y, "b2c_bn1d_apply: null pointer");
  B2C_BN1D_CHECK("b2c_bn1d_apply");
  const int64_t n8 = rows * (C / 8);
  const int grid = static_cast<int>(n8 / 256 + 1 < 8 * 148 ? n8 / 256 + 1 : 8 * 148);
  B2C_CUDA(launch_k(bn1d_apply_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(h), mean, rstd,
                                                                         gamma, beta, y, rows, C));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_bn1d_bwd_reduce(const float* dy, const void* h, const float* mean, const float* rstd, double* sums,
                                   int64_t rows, int C, void* stream) {
  B2C_REQUIRE(dy && h && mean && rstd && sums, "b2c_bn1d_bwd_reduce: null pointer");
  B2C_BN1D_CHECK("b2c_bn1d_bwd_reduce");
  int64_t rpb;
  const int grid = bn1d_grid(rows, C, &rpb);
  B2C_CUDA(launch_k(bn1d_bwd_reduce_kernel, dim3(grid), dim3(256), 2 * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      dy, static_cast<const __nv_bfloat16*>(h), mean, rstd, sums, rows, C, static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_bn1d_bwd_apply(const float* dy, const void* h, const void* z, const float* mean, const float* rstd,
                                  const float* gamma, const double* sums, int train, void* dz, float* colsum, int64_t rows,
                                  int C, void* stream) {
  B2C_REQUIRE(dy && h && z && mean && rstd && gamma && sums && dz, "b2c_bn1d_bwd_apply: null pointer");
  B2C_BN1D_CHECK("b2c_bn1d_bwd_apply");
  int64_t rpb;
  const int grid = bn1d_grid(rows, C, &rpb);
  B2C_CUDA(launch_k(bn1d_bwd_apply_kernel, dim3(grid), dim3(256), C * sizeof(float), static_cast<cudaStream_t>(stream), 
      dy, static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma, sums,
      1.0 / static_cast<double>(rows), train, static_cast<__nv_bfloat16*>(dz), colsum, rows, C, static_cast<int>(rpb)));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
