// Fused multi-tensor Adafactor step (the optimizer of the reference's training loop:
// transformers.Adafactor(lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False), /root/reference/lib/standard/runner.py:36,146).
//
// HF's implementation walks the parameters in Python and launches ~15 small kernels per tensor (~7000 launches per
// step for Conformer-M).  Here the whole parameter set is updated by FOUR launches driven by device-side tables:
//   1. adafactor_stats   : row / column sums of (g^2 + eps1) for factored (>= 2-D) tensors
//   2. adafactor_factors : exp_avg_sq_row/col <- beta2t * . + (1-beta2t) * mean ; per-slab mean of the row factor
//   3. adafactor_update  : u = g * rsqrt(row/mean(row)) * rsqrt(col)   (1-D: v <- beta2t v + (1-beta2t)(g^2+eps1), u = g rsqrt(v))
//                          and sum(u^2) per tensor (for the RMS clip)
//   4. adafactor_apply   : u /= max(1, rms(u)/clip); m <- beta1 m + (1-beta1) lr u; p -= m
// Every tensor is viewed as [P slabs][R rows][C cols] (P = prod(shape[:-2]), 1-D tensors as [1][1][n]); HF's
// exp_avg_sq_row has shape [P,R] (mean over the last dim), exp_avg_sq_col [P,C] (mean over dim -2).
// Memory-bound: g is read 3x, m and p read+written once (28 B per parameter).
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

struct AfTensor {            // one per parameter (device table, 96 bytes)
  float* p;                  // parameter
  const float* g;            // gradient
  float* m;                  // exp_avg
  float* row;                // exp_avg_sq_row [P*R]  (factored)  | exp_avg_sq [n] (1-D)
  float* col;                // exp_avg_sq_col [P*C]  (factored)
  float* rowsum;             // workspace [P*R]
  float* colsum;             // workspace [P*C]
  float* rmean;              // workspace [P]   mean over R of the updated row factor
  float* sumsq;              // workspace [1]   sum(u^2)
  int32_t P, R, C, factored;
  int64_t n;
};
struct AfChunk {             // a run of `nrows` flattened rows (row index in [0, P*R)) of tensor `t`
  int32_t t, row0, nrows, pad;
};

constexpr int AF_THREADS = 256;

__global__ void __launch_bounds__(AF_THREADS) adafactor_stats_kernel(const AfTensor* __restrict__ tensors,
                                                                    const AfChunk* __restrict__ chunks, float eps1) {
  pdl_trigger();
  pdl_wait();
  const AfChunk ck = chunks[blockIdx.x];
  const AfTensor T = tensors[ck.t];
  if (!T.factored) return;
  const int C = T.C, R = T.R;
  const float* gbase = T.g + static_cast<int64_t>(ck.row0) * C;
  if ((C & 127) == 0 && C <= 4 * AF_THREADS && (reinterpret_cast<uintptr_t>(gbase) & 15) == 0 &&
      ck.row0 / R == (ck.row0 + ck.nrows - 1) / R) {
    // the linear weights (C = 256 / 1024, one slab per chunk): ONE pass over g with 16-byte loads.  A thread owns four
    // adjacent columns of every RPI-th row; a row's C/4 threads are whole warps, so the row sum is a shuffle reduction
    // plus one atomic per warp (rowsum is zeroed with the rest of the workspace), the column sums stay in registers.
    const int tid = threadIdx.x;
    const int G = C >> 2, RPI = AF_THREADS / G;
    const int tr = tid / G, tc = tid - tr * G;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tr < RPI) {
      const float4* gp = reinterpret_cast<const float4*>(gbase) + tc;
      const bool lane0 = (tid & 31) == 0;
      constexpr int UN = 4;
      for (int r = tr; r < ck.nrows; r += UN * RPI) {
        float4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int ru = r + u * RPI;
          v[u] = ru < ck.nrows ? gp[static_cast<int64_t>(ru) * G] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int ru = r + u * RPI;
          if (ru < ck.nrows) {                          // uniform over the warp
            const float sx = fmaf(v[u].x, v[u].x, eps1), sy = fmaf(v[u].y, v[u].y, eps1);
            const float sz = fmaf(v[u].z, v[u].z, eps1), sw = fmaf(v[u].w, v[u].w, eps1);
            acc.x += sx; acc.y += sy; acc.z += sz; acc.w += sw;
            const float rs = warp_sum((sx + sy) + (sz + sw));
            if (lane0) atomicAdd(T.rowsum + ck.row0 + ru, rs);
          }
        }
      }
    }
    __shared__ float4 cs[AF_THREADS];
    cs[tid] = acc;
    __syncthreads();
    if (tid < G) {
      float4 t = cs[tid];
      for (int k = 1; k < RPI; ++k) {
        const float4 o = cs[tid + k * G];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      float* dst = T.colsum + static_cast<int64_t>(ck.row0 / R) * C + 4 * tid;
      atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
    }
  } else if (C >= 32) {
    // wide rows: warp w takes rows w, w+8, ...; lanes stride the columns.  Row sums by warp shuffle; column sums are
    // kept per (warp, column) in registers only for C <= 32*8*... -> simpler: per-element atomics on the column sums
    // are avoided by letting each THREAD own columns c = tid, tid+256, ... across all rows of the chunk.
    const int tid = threadIdx.x;
    for (int c = tid; c < C; c += AF_THREADS) {
      float acc = 0.f;
      int cur_p = -1;
      for (int r = 0; r < ck.nrows; ++r) {
        const int row = ck.row0 + r;
        const int pslab = row / R;
        if (pslab != cur_p) {
          if (cur_p >= 0) atomicAdd(T.colsum + static_cast<int64_t>(cur_p) * C + c, acc);
          acc = 0.f;
          cur_p = pslab;
        }
        const float gv = T.g[static_cast<int64_t>(row) * C + c];
        acc += gv * gv + eps1;
      }
      if (cur_p >= 0) atomicAdd(T.colsum + static_cast<int64_t>(cur_p) * C + c, acc);
    }
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < ck.nrows; r += AF_THREADS / 32) {
      const int row = ck.row0 + r;
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float gv = T.g[static_cast<int64_t>(row) * C + c];
        acc += gv * gv + eps1;
      }
      acc = warp_sum(acc);
      if (lane == 0) T.rowsum[row] = acc;           // a row belongs to exactly one chunk: plain store
    }
  } else {
    // narrow rows (conv kernels, C = 1..31): one thread per row.  When the 32 rows of a warp lie in one slab (pointwise
    // convs: 256 rows of one column) the column sums are reduced in the warp first -- one atomic per warp instead of 32
    // on the same address.
    const int lane = threadIdx.x & 31;
    for (int r0 = 0; r0 < ck.nrows; r0 += AF_THREADS) {
      const int r = r0 + static_cast<int>(threadIdx.x);
      const bool ok = r < ck.nrows;
      const int row = ck.row0 + (ok ? r : ck.nrows - 1);
      const int pslab = row / R;
      const int p0 = __shfl_sync(0xffffffffu, pslab, 0);
      const bool uni = __all_sync(0xffffffffu, pslab == p0);
      float acc = 0.f;
      for (int c = 0; c < C; ++c) {
        const float gv = T.g[static_cast<int64_t>(row) * C + c];
        const float s = ok ? gv * gv + eps1 : 0.f;
        acc += s;
        if (uni) {
          const float t = warp_sum(s);
          if (lane == 0) atomicAdd(T.colsum + static_cast<int64_t>(p0) * C + c, t);
        } else if (ok) {
          atomicAdd(T.colsum + static_cast<int64_t>(pslab) * C + c, s);
        }
      }
      if (ok) T.rowsum[row] = acc;
    }
  }
}

// step counter and beta2_t live on the device so a captured CUDA graph advances them on every replay
__global__ void adafactor_tick_kernel(int64_t* __restrict__ step, float* __restrict__ beta2t, float decay_rate) {
  pdl_trigger();
  pdl_wait();
  const int64_t s = *step + 1;
  *step = s;
  *beta2t = 1.f - powf(static_cast<float>(s), decay_rate);          // HF: 1 - step^decay_rate
}

// grid = (blocks, tensors): updates the factored second moments and the per-slab mean of the row factor.
// A warp owns whole slabs (it updates the R row entries of slab p and reduces their mean itself), columns are a
// plain grid-stride loop -- no cross-block dependency, so large stacks of small slabs (conv kernels: 65k slabs of 3x3)
// spread over the whole chip instead of crawling through one CTA.
__global__ void __launch_bounds__(AF_THREADS) adafactor_factors_kernel(const AfTensor* __restrict__ tensors, int n_tensors,
                                                                      const float* __restrict__ beta2t_ptr) {
  pdl_trigger();
  pdl_wait();
  const float beta2t = *beta2t_ptr;
  const AfTensor T = tensors[blockIdx.y];
  if (!T.factored) return;
  const int C = T.C, R = T.R, P = T.P;
  const float om = 1.f - beta2t;
  const int64_t gtid = static_cast<int64_t>(blockIdx.x) * AF_THREADS + threadIdx.x;
  const int64_t gstride = static_cast<int64_t>(gridDim.x) * AF_THREADS;
  for (int64_t i = gtid; i < static_cast<int64_t>(P) * C; i += gstride)
    T.col[i] = beta2t * T.col[i] + om * (T.colsum[i] / R);
  if (R <= 8) {                                        // stacks of tiny slabs (3x3 / 7x7 conv kernels): a thread per slab
    for (int64_t pslab = gtid; pslab < P; pslab += gstride) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) {
        const int64_t i = pslab * R + r;
        const float v = beta2t * T.row[i] + om * (T.rowsum[i] / C);
        T.row[i] = v;
        acc += v;
      }
      T.rmean[pslab] = acc / R;
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = gtid >> 5, nwarps = gstride >> 5;
  for (int64_t pslab = gwarp; pslab < P; pslab += nwarps) {
    float acc = 0.f;
    for (int r = lane; r < R; r += 32) {
      const int64_t i = pslab * R + r;
      const float v = beta2t * T.row[i] + om * (T.rowsum[i] / C);
      T.row[i] = v;
      acc += v;
    }
    acc = warp_sum(acc);
    if (lane == 0) T.rmean[pslab] = acc / R;
  }
}

template <bool APPLY>
__global__ void __launch_bounds__(AF_THREADS) adafactor_update_kernel(const AfTensor* __restrict__ tensors,
                                                                     const AfChunk* __restrict__ chunks,
                                                                     const float* __restrict__ beta2t_ptr, float eps1,
                                                                     float lr, float beta1, float clip) {
  pdl_trigger();
  pdl_wait();
  const float beta2t = *beta2t_ptr;
  const AfChunk ck = chunks[blockIdx.x];
  const AfTensor T = tensors[ck.t];
  const int C = T.C, R = T.R;
  const int64_t e0 = static_cast<int64_t>(ck.row0) * C;
  const int64_t cnt = static_cast<int64_t>(ck.nrows) * C;
  float scale = 0.f;
  if (APPLY) {
    const float rms = sqrtf(*T.sumsq / static_cast<float>(T.n));
    scale = lr / fmaxf(1.f, rms / clip);
  }
  float ss = 0.f;
  const int icnt = static_cast<int>(cnt);
  // 16-byte path: four elements of one row per thread and step (every large tensor has C % 4 == 0 and 16-byte aligned
  // slices); the scalar loop below remains for the small odd-shaped ones
  const bool vec = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(T.g + e0) | reinterpret_cast<uintptr_t>(T.p + e0) |
                                     reinterpret_cast<uintptr_t>(T.m + e0)) & 15) == 0 &&
                   (T.factored ? (reinterpret_cast<uintptr_t>(T.col) & 15) == 0 : (reinterpret_cast<uintptr_t>(T.row + e0) & 15) == 0);
  if (vec) {
    const int step = AF_THREADS * 4;
    const int vstep_r = step / C, vstep_c = step % C;
    int vr = static_cast<int>(threadIdx.x * 4) / C, vc = static_cast<int>(threadIdx.x * 4) % C;
    for (int i = threadIdx.x * 4; i < icnt; i += step) {
      const int64_t e = e0 + i;
      const float4 g4 = *reinterpret_cast<const float4*>(T.g + e);
      float u[4];
      if (T.factored) {
        const int row = ck.row0 + vr;
        const int pslab = R == 1 ? row : row / R;
        const float rs = rsqrtf(T.row[row] / T.rmean[pslab]);
        const float4 c4 = *reinterpret_cast<const float4*>(T.col + static_cast<int64_t>(pslab) * C + vc);
        u[0] = g4.x * rs * rsqrtf(c4.x); u[1] = g4.y * rs * rsqrtf(c4.y);
        u[2] = g4.z * rs * rsqrtf(c4.z); u[3] = g4.w * rs * rsqrtf(c4.w);
      } else {
        float4 v4 = *reinterpret_cast<const float4*>(T.row + e);
        if (!APPLY) {                                // second moment is updated once, in the first pass
          v4.x = beta2t * v4.x + (1.f - beta2t) * (g4.x * g4.x + eps1); v4.y = beta2t * v4.y + (1.f - beta2t) * (g4.y * g4.y + eps1);
          v4.z = beta2t * v4.z + (1.f - beta2t) * (g4.z * g4.z + eps1); v4.w = beta2t * v4.w + (1.f - beta2t) * (g4.w * g4.w + eps1);
          *reinterpret_cast<float4*>(T.row + e) = v4;
        }
        u[0] = g4.x * rsqrtf(v4.x); u[1] = g4.y * rsqrtf(v4.y); u[2] = g4.z * rsqrtf(v4.z); u[3] = g4.w * rsqrtf(v4.w);
      }
      if (APPLY) {
        float4 m4 = *reinterpret_cast<const float4*>(T.m + e);
        float4 p4 = *reinterpret_cast<const float4*>(T.p + e);
        m4.x = beta1 * m4.x + (1.f - beta1) * (scale * u[0]); m4.y = beta1 * m4.y + (1.f - beta1) * (scale * u[1]);
        m4.z = beta1 * m4.z + (1.f - beta1) * (scale * u[2]); m4.w = beta1 * m4.w + (1.f - beta1) * (scale * u[3]);
        p4.x -= m4.x; p4.y -= m4.y; p4.z -= m4.z; p4.w -= m4.w;
        *reinterpret_cast<float4*>(T.m + e) = m4;
        *reinterpret_cast<float4*>(T.p + e) = p4;
      } else {
        ss += u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3];
      }
      vr += vstep_r; vc += vstep_c;
      if (vc >= C) { vc -= C; ++vr; }
    }
  }
  // (row, col) of this thread's element advance incrementally: no per-element 64-bit division
  if (!vec) {
  const int step_r = AF_THREADS / C, step_c = AF_THREADS % C;
  int rr = static_cast<int>(threadIdx.x) / C, c = static_cast<int>(threadIdx.x) % C;
  for (int i = threadIdx.x; i < icnt; i += AF_THREADS) {
    const int64_t e = e0 + i;
    const float gv = T.g[e];
    float u;
    if (T.factored) {
      const int row = ck.row0 + rr;
      const int pslab = R == 1 ? row : row / R;
      u = gv * rsqrtf(T.row[row] / T.rmean[pslab]) * rsqrtf(T.col[static_cast<int64_t>(pslab) * C + c]);
    } else {
      float v = T.row[e];
      if (!APPLY) {                                  // second moment is updated once, in the first pass
        v = beta2t * v + (1.f - beta2t) * (gv * gv + eps1);
        T.row[e] = v;
      }
      u = gv * rsqrtf(v);
    }
    if (APPLY) {
      const float mv = beta1 * T.m[e] + (1.f - beta1) * (scale * u);
      T.m[e] = mv;
      T.p[e] -= mv;
    } else {
      ss += u * u;
    }
    rr += step_r; c += step_c;
    if (c >= C) { c -= C; ++rr; }
  }
  }
  if (!APPLY) {
    __shared__ float red[AF_THREADS / 32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < AF_THREADS / 32; ++w) t += red[w];
      atomicAdd(T.sumsq, t);
    }
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_adafactor_step(const void* tensors_dev, int n_tensors, const void* chunks_dev, int n_chunks,
                                  int64_t* step_dev, float* beta2t_dev, float lr, float beta1, float eps1, float clip,
                                  float decay_rate, void* stream) {
  B2C_REQUIRE(tensors_dev && chunks_dev && step_dev && beta2t_dev && n_tensors > 0 && n_chunks > 0,
              "b2c_adafactor_step: bad argument");
  static_assert(sizeof(AfTensor) == 96, "AfTensor layout is part of the ABI (python builds the table)");
  static_assert(sizeof(AfChunk) == 16, "AfChunk layout is part of the ABI");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AfTensor* T = static_cast<const AfTensor*>(tensors_dev);
  const AfChunk* Cc = static_cast<const AfChunk*>(chunks_dev);
  B2C_CUDA(launch_k(adafactor_tick_kernel, dim3(1), dim3(1), 0, st, step_dev, beta2t_dev, decay_rate));
  B2C_CUDA(launch_k(adafactor_stats_kernel, dim3(n_chunks), dim3(AF_THREADS), 0, st, T, Cc, eps1));
  B2C_CUDA(launch_k(adafactor_factors_kernel, dim3(dim3(32, n_tensors)), dim3(AF_THREADS), 0, st, T, n_tensors, beta2t_dev));
  B2C_CUDA(launch_k(adafactor_update_kernel<false>, dim3(n_chunks), dim3(AF_THREADS), 0, st, T, Cc, beta2t_dev, eps1, lr, beta1, clip));
  B2C_CUDA(launch_k(adafactor_update_kernel<true>, dim3(n_chunks), dim3(AF_THREADS), 0, st, T, Cc, beta2t_dev, eps1, lr, beta1, clip));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
