// ConvSubSampling front end (/root/reference/lib/convsubsampling.py:16-45) on tcgen05.
//
// The reference applies Conv2d(1,C1,k1,stride s1) then Conv2d(C1,C2,k2,stride s2) with NOTHING in between
// (convsubsampling.py:43-44), so the pair is one linear map: Conv2d(1, C2, kc x kc, stride s1*s2) with
// kc = k1 + (k2-1)*s1, Wc[o,s1*u+p,s1*v+q] += W2[o,c,u,v]*W1[c,p,q], bc = b2 + sum W2*b1 (SURVEY.md Appendix A,
// checked to 1e-13 against the reference in tests/test_oracle_cpu.py).  For the defaults (C1=512,k1=7,k2=3,s=2)
// that is an 11x11 stride-4 conv: 45x fewer FLOPs and no (B,512,37,T1) intermediate (1.8 GB at 64 x 15 s).
// The parameters stay the module's own conv_sub_1 / conv_sub_2 tensors; gradients flow back through the
// composition (convsub_compose_bwd), so training updates exactly the reference's parameters.
//
// forward  : implicit GEMM.  CTA = (128 output frames t', one mel row f', utterance b).  The 128 x KP im2col tile is
//            gathered straight into the K-major SWIZZLE_128B UMMA layout (lanes = consecutive t' -> 16 B stride,
//            4 cache lines per warp load), Wc arrives by TMA, 4*KP/64 tcgen05.mma (M128, N=C2, K16), epilogue adds
//            bc and writes token-major rows out[(b*T'+t')*F' + f'][C2] (bf16, or fp32 for the module API).
// wgrad    : dWc[c,kk] = sum_m dout[m,c]*patch[m,kk]: the SAME smem im2col tile read as an MN-major B operand,
//            dout tiles by TMA as MN-major A; CTAs stride over tiles, accumulate in TMEM, one fp32 atomic pass.
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

struct ConvSubParams {
  const float* x;
  const float* bc;
  void* out;
  float* dWc;
  int out_f32;
  int B, F, T, Fo, To, C2, kc, stride, KP;
  int tiles_t, total_tiles;
};

// im2col: rows = 128 consecutive output frames starting at to0 for (b, fo); K index kk = p*kc + q -> x[b, fo*s+p, t*s+q]
// written as bf16 into KP/64 K-major SW128 blocks of [128 rows][128 B].  Rows with t' >= To and kk >= kc*kc are zero.
__device__ __forceinline__ void im2col_tile(const ConvSubParams& p, uint8_t* sA, int b, int fo, int to0) {
  const int row = threadIdx.x & 127;
  const int par = threadIdx.x >> 7;                 // 256 threads: two chunk phases
  const int to = to0 + row;
  const bool row_ok = to < p.To;
  const float* xb = p.x + (static_cast<int64_t>(b) * p.F + fo * p.stride) * p.T + to * p.stride;
  const int kk_end = p.kc * p.kc;
  for (int chunk = par; chunk < p.KP / 8; chunk += 2) {
    float f[8];
    int kk = chunk * 8;
    int pr = kk / p.kc, q = kk - pr * p.kc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = (row_ok && kk + j < kk_end) ? __ldg(xb + static_cast<int64_t>(pr) * p.T + q) : 0.f;
      if (++q == p.kc) { q = 0; ++pr; }
    }
    const int blk = chunk >> 3, cc = chunk & 7;
    const uint32_t off = static_cast<uint32_t>(blk * 16384 + row * 128 + ((cc ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(sA + off) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

__device__ __forceinline__ uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (static_cast<int>(c) < n) c <<= 1;
  return c;
}

__global__ void __launch_bounds__(256) convsub_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const ConvSubParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int KB = p.KP / 64;
  uint8_t* sA = smem;                               // KB x 16 KB
  const uint32_t w_off = KB * 16384;
  const uint32_t w_blk = static_cast<uint32_t>(p.C2) * 128;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + w_off + KB * w_blk);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int to0 = blockIdx.x * 128, fo = blockIdx.y, b = blockIdx.z;
  const uint32_t ncols = pow2_cols(p.C2);

  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, ncols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_w, KB * w_blk);
    for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem + w_off + kb * w_blk, &tmW, bar_w, kb * 64, 0);
  }
  im2col_tile(p, sA, b, fo, to0);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(bar_w, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.C2, false, false);
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, umma_desc_sw128(base + kb * 16384 + k * 32, 16, 1024),
                  umma_desc_sw128(base + w_off + kb * w_blk + k * 32, 16, 1024), idesc, (kb | k) ? 1u : 0u);
    tc_commit(bar_mma);
  }
  if (warp < 4) {                                   // warps 0-3 <-> TMEM lane quarters 0-3
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    const int to = to0 + warp * 32 + lane;
    const int64_t orow = (static_cast<int64_t>(b) * p.To + to) * p.Fo + fo;
    for (int c = 0; c < p.C2; c += 16) {
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_wait_ld();
      if (to < p.To) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __ldg(p.bc + c + j);
        if (p.out_f32) {
          float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.C2 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.C2 + c);
          o[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          o[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, ncols);
  }
}

// dWc[c, kk] += sum over this CTA's tiles of dout[m, c] * patch[m, kk]   (M-dim = c, N-dim = kk, reduction = frames)
__global__ void __launch_bounds__(256) convsub_wgrad_kernel(const __grid_constant__ CUtensorMap tmD, const ConvSubParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int KB = p.KP / 64;
  uint8_t* sP = smem;                               // im2col tile, KB chunks x [128 frames][128 B]  (B operand, MN-major)
  const uint32_t d_off = KB * 16384;
  constexpr int MC = 2;                             // this CTA's 128 output channels = 2 x 64-wide chunks of dout
  const int half = blockIdx.y;                      // which 128-channel slab of C2
  uint64_t* bar_d = reinterpret_cast<uint64_t*>(smem + d_off + MC * 16384);
  uint64_t* bar_mma = bar_d + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncols = pow2_cols(p.KP);

  if (tid == 0) {
    tma_prefetch_desc(&tmD);
    mbar_init(bar_d, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, ncols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, p.KP, true, true);

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int tb = tile % p.tiles_t;
    const int fo = (tile / p.tiles_t) % p.Fo;
    const int b = tile / (p.tiles_t * p.Fo);
    const int to0 = tb * 128;
    if (tid == 0) {
      mbar_expect_tx(bar_d, MC * 16384);
      for (int c = 0; c < MC; ++c)
        tma_load_2d(smem + d_off + c * 16384, &tmD, bar_d, fo * p.C2 + half * 128 + c * 64, b * p.To + to0);
    }
    im2col_tile(p, sP, b, fo, to0);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_d, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < 8; ++k)                   // 128 frames = 8 x K16
        umma_bf16(tmem_base, umma_desc_sw128(base + d_off + k * 2048, 16384, 1024),
                  umma_desc_sw128(base + k * 2048, 16384, 1024), idesc, (it | k) ? 1u : 0u);
      tc_commit(bar_mma);
      mbar_wait(bar_mma, it & 1);                   // smem tiles are reused by the next iteration
    }
    __syncthreads();
  }
  tc_fence_after();
  if (warp < 4 && it > 0) {
    const int c = half * 128 + warp * 32 + lane;    // accumulator row = output channel
    for (int n = 0; n < p.KP; n += 16) {
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + n, v);
      tmem_wait_ld();
      if (c < p.C2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(p.dWc + static_cast<int64_t>(c) * p.KP + n + j, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ----------------------------------------------------------------------------- weight composition (tiny)
// One output element per group of CS_SPLIT lanes, which split the C1-long channel sum and meet in a shuffle reduction: with
// one THREAD per element the whole composition was 16 k threads running 4608 dependent FMAs each (221 us).
constexpr int CS_SPLIT = 32;     // a whole warp per output element: 16 channels per lane (8 lanes: 62 us; 32 lanes measured below)
__global__ void __launch_bounds__(256) compose_w_kernel(const float* __restrict__ W1, const float* __restrict__ W2, int C1,
                                                        int C2, int k1, int s1, int k2, int kc, int KP,
                                                        __nv_bfloat16* __restrict__ Wc, float* __restrict__ Wc_f32) {
  pdl_trigger();
  pdl_wait();
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gidx / CS_SPLIT, part = gidx % CS_SPLIT;    // (o, kk) over C2 x KP; part = which slice of the channels
  const bool live = idx < C2 * KP;
  const int o = live ? idx / KP : 0, kk = live ? idx % KP : 0;
  float acc = 0.f;
  if (live && kk < kc * kc) {
    const int y = kk / kc, x = kk % kc;
    for (int u = 0; u < k2; ++u) {
      const int pp = y - s1 * u;
      if (pp < 0 || pp >= k1) continue;
      for (int v = 0; v < k2; ++v) {
        const int q = x - s1 * v;
        if (q < 0 || q >= k1) continue;
        const float* w2 = W2 + (static_cast<int64_t>(o) * C1 * k2 + u) * k2 + v;
        const float* w1 = W1 + pp * k1 + q;
        for (int c = part; c < C1; c += CS_SPLIT)
          acc = fmaf(w2[static_cast<int64_t>(c) * k2 * k2], w1[static_cast<int64_t>(c) * k1 * k1], acc);
      }
    }
  }
#pragma unroll
  for (int off = CS_SPLIT / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (live && part == 0) {
    Wc[idx] = __float2bfloat16(acc);
    if (Wc_f32) Wc_f32[idx] = acc;
  }
}

__global__ void compose_b_kernel(const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                                 int C1, int C2, int k2, float* __restrict__ bc) {
  pdl_trigger();
  pdl_wait();
  const int o = blockIdx.x, lane = threadIdx.x;               // one warp per output channel
  float acc = 0.f;
  const int n = C1 * k2 * k2;
  for (int i = lane; i < n; i += 32) acc = fmaf(W2[static_cast<int64_t>(o) * n + i], b1[i / (k2 * k2)], acc);
  acc = warp_sum(acc);
  if (lane == 0) bc[o] = acc + b2[o];
}

__global__ void __launch_bounds__(256) compose_bwd_w1_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc,
                                                             const float* __restrict__ W2, int C1, int C2, int k1, int s1,
                                                             int k2, int kc, int KP, float* __restrict__ dW1,
                                                             float* __restrict__ db1) {
  pdl_trigger();
  pdl_wait();
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gidx / CS_SPLIT, part = gidx % CS_SPLIT;    // (c, p, q) plus one extra slot per c for db1; part = slice of o
  const int per_c = k1 * k1 + 1;
  const bool live = idx < C1 * per_c;
  const int c = live ? idx / per_c : 0, r = live ? idx % per_c : 0;
  float acc = 0.f;
  if (live && r == k1 * k1) {
    for (int o = part; o < C2; o += CS_SPLIT)
      for (int uv = 0; uv < k2 * k2; ++uv) acc = fmaf(dbc[o], W2[(static_cast<int64_t>(o) * C1 + c) * k2 * k2 + uv], acc);
  } else if (live) {
    const int pp = r / k1, q = r % k1;
    for (int o = part; o < C2; o += CS_SPLIT)
      for (int u = 0; u < k2; ++u)
        for (int v = 0; v < k2; ++v)
          acc = fmaf(dWc[static_cast<int64_t>(o) * KP + (s1 * u + pp) * kc + (s1 * v + q)],
                     W2[((static_cast<int64_t>(o) * C1 + c) * k2 + u) * k2 + v], acc);
  }
#pragma unroll
  for (int off = CS_SPLIT / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (live && part == 0) {
    if (r == k1 * k1) db1[c] += acc;
    else dW1[idx - c] += acc;                                  // dW1 index = c*k1*k1 + r
  }
}

// ---- staged versions (used whenever their shared-memory plan fits; the generic kernels above stay as the fallback) ----
// compose: one CTA per output channel o.  W2[o] (C1 x k2 x k2) and ALL of W1 (C1 x k1 x k1) are staged in shared memory with
// coalesced loads; thread = (output tap kk, half of the channel sum): W2 reads are warp-wide broadcasts, W1 reads walk along a
// row.  Also produces bc[o] (compose_b_kernel's job).  The warp-per-element kernel above spent its time in 36..196-byte
// strided global loads (one useful word per 32-byte sector): 54 + 29 us against ~10 us for this one at 512 -> 128 channels.
__global__ void __launch_bounds__(256) compose_staged_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                                             const float* __restrict__ W2, const float* __restrict__ b2, int C1,
                                                             int C2, int k1, int s1, int k2, int kc, int KP,
                                                             __nv_bfloat16* __restrict__ Wc, float* __restrict__ Wc_f32,
                                                             float* __restrict__ bc) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float cs_sm[];
  const int n2 = C1 * k2 * k2, n1 = C1 * k1 * k1;
  float* sW2 = cs_sm;                    // [C1][k2*k2]   this o's slice of W2
  float* sW1 = cs_sm + n2;               // [C1][k1*k1]
  float* sred = sW1 + n1;                // [256] partial sums
  const int o = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < n2; i += 256) sW2[i] = W2[static_cast<int64_t>(o) * n2 + i];
  for (int i = tid; i < n1; i += 256) sW1[i] = W1[i];
  __syncthreads();
  const int half = tid >> 7, kl = tid & 127;
  const int c_lo = half ? C1 / 2 : 0, c_hi = half ? C1 : C1 / 2;
  for (int kk0 = 0; kk0 < KP; kk0 += 128) {
    const int kk = kk0 + kl;
    float acc = 0.f;
    if (kk < kc * kc) {
      const int y = kk / kc, x = kk % kc;
      for (int u = 0; u < k2; ++u) {
        const int pp = y - s1 * u;
        if (pp < 0 || pp >= k1) continue;
        for (int v = 0; v < k2; ++v) {
          const int q = x - s1 * v;
          if (q < 0 || q >= k1) continue;
          const float* w2 = sW2 + u * k2 + v;
          const float* w1 = sW1 + pp * k1 + q;
          float a0 = 0.f, a1 = 0.f;
          int c = c_lo;
          for (; c + 1 < c_hi; c += 2) {
            a0 = fmaf(w2[c * k2 * k2], w1[c * k1 * k1], a0);
            a1 = fmaf(w2[(c + 1) * k2 * k2], w1[(c + 1) * k1 * k1], a1);
          }
          if (c < c_hi) a0 = fmaf(w2[c * k2 * k2], w1[c * k1 * k1], a0);
          acc += a0 + a1;
        }
      }
    }
    sred[tid] = acc;
    __syncthreads();
    if (half == 0 && kk < KP) {
      const float tot = acc + sred[tid + 128];
      Wc[static_cast<int64_t>(o) * KP + kk] = __float2bfloat16(tot);
      if (Wc_f32) Wc_f32[static_cast<int64_t>(o) * KP + kk] = tot;
    }
    __syncthreads();
  }
  // bc[o] = b2[o] + sum_c b1[c] * sum_uv W2[o, c, u, v]
  float part = 0.f;
  for (int i = tid; i < n2; i += 256) part = fmaf(sW2[i], b1[i / (k2 * k2)], part);
  part = warp_sum(part);
  if ((tid & 31) == 0) sred[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float tot = b2[o];
    for (int w = 0; w < 8; ++w) tot += sred[w];
    bc[o] = tot;
  }
}

// backward of the composition w.r.t. W1 / b1: thread = one element (c, p, q) (+ one slot per c for db1), blockIdx.y = slice of
// (at most 32) output channels o whose dWc rows are staged in shared memory; W2[o, c, :, :] is 9 contiguous floats and the same
// address for (almost) all lanes of a warp.  Slices meet in dW1 / db1 through atomics.  The warp-per-element kernel above read
// one useful word per 32-byte sector of W2 (18 KB lane stride): 212 us against ~10 us at 512 -> 128 channels.
constexpr int CB_OSLICE = 32;
__global__ void __launch_bounds__(256) compose_bwd_w1_staged_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc,
                                                                    const float* __restrict__ W2, int C1, int C2, int k1, int s1,
                                                                    int k2, int kc, int KP, float* __restrict__ dW1,
                                                                    float* __restrict__ db1) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float cb_sm[];        // [CB_OSLICE][kc*kc] dWc rows, then [CB_OSLICE] dbc
  const int kk2 = kc * kc;
  const int o0 = blockIdx.y * CB_OSLICE, no = min(CB_OSLICE, C2 - o0);
  float* s_dbc = cb_sm + CB_OSLICE * kk2;
  for (int i = threadIdx.x; i < no * kk2; i += 256) cb_sm[i] = dWc[static_cast<int64_t>(o0 + i / kk2) * KP + i % kk2];
  for (int i = threadIdx.x; i < no; i += 256) s_dbc[i] = dbc[o0 + i];
  __syncthreads();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int per_c = k1 * k1 + 1;
  if (idx >= C1 * per_c) return;
  const int c = idx / per_c, r = idx % per_c;
  const int kw = k2 * k2;
  float acc = 0.f;
  if (r == k1 * k1) {
    for (int o = 0; o < no; ++o) {
      const float* w = W2 + (static_cast<int64_t>(o0 + o) * C1 + c) * kw;
      float sw = 0.f;
      for (int uv = 0; uv < kw; ++uv) sw += w[uv];
      acc = fmaf(s_dbc[o], sw, acc);
    }
    atomicAdd(db1 + c, acc);
  } else {
    const int pp = r / k1, q = r % k1;
    for (int o = 0; o < no; ++o) {
      const float* w = W2 + (static_cast<int64_t>(o0 + o) * C1 + c) * kw;
      const float* dw = cb_sm + o * kk2 + pp * kc + q;
      for (int u = 0; u < k2; ++u)
        for (int v = 0; v < k2; ++v) acc = fmaf(dw[(s1 * u) * kc + s1 * v], w[u * k2 + v], acc);
    }
    atomicAdd(dW1 + static_cast<int64_t>(c) * k1 * k1 + r, acc);
  }
}

__global__ void compose_bwd_w2_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc,
                                      const float* __restrict__ W1, const float* __restrict__ b1, int C1, int C2, int k1,
                                      int s1, int k2, int kc, int KP, float* __restrict__ dW2, float* __restrict__ db2) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (o, c, u, v)
  if (idx >= static_cast<int64_t>(C2) * C1 * k2 * k2) return;
  const int v = idx % k2, u = (idx / k2) % k2;
  const int c = (idx / (k2 * k2)) % C1, o = static_cast<int>(idx / (static_cast<int64_t>(k2) * k2 * C1));
  float acc = dbc[o] * b1[c];                                 // bc = b2 + sum W2*b1
  for (int pp = 0; pp < k1; ++pp)
    for (int q = 0; q < k1; ++q)
      acc = fmaf(dWc[static_cast<int64_t>(o) * KP + (s1 * u + pp) * kc + (s1 * v + q)], W1[(c * k1 + pp) * k1 + q], acc);
  dW2[idx] += acc;
  if (c == 0 && u == 0 && v == 0) db2[o] += dbc[o];
}

static int convsub_geometry(ConvSubParams& p, int B, int F, int T, int C2, int kc, int stride, int KP, const char* who) {
  B2C_REQUIRE(B > 0 && F >= kc && T >= kc, "%s: input %dx%d smaller than the composed kernel %d", who, F, T, kc);
  B2C_REQUIRE(KP % 64 == 0 && KP >= kc * kc && KP <= 256, "%s: KP=%d must be a multiple of 64 in [kc*kc, 256]", who, KP);
  B2C_REQUIRE(C2 % 16 == 0 && C2 >= 16 && C2 <= 256, "%s: out channels %d must be a multiple of 16, <= 256", who, C2);
  p.B = B; p.F = F; p.T = T; p.C2 = C2; p.kc = kc; p.stride = stride; p.KP = KP;
  p.Fo = (F - kc) / stride + 1;
  p.To = (T - kc) / stride + 1;
  p.tiles_t = (p.To + 127) / 128;
  p.total_tiles = p.tiles_t * p.Fo * B;
  return 0;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_convsub_compose(const float* W1, const float* b1, const float* W2, const float* b2, int C1, int C2,
                                   int k1, int s1, int k2, void* Wc, float* Wc_f32, float* bc, int KP, void* stream) {
  B2C_REQUIRE(W1 && b1 && W2 && b2 && Wc && bc, "b2c_convsub_compose: null pointer");
  const int kc = k1 + (k2 - 1) * s1;
  B2C_REQUIRE(KP >= kc * kc, "b2c_convsub_compose: KP too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem_staged = (static_cast<size_t>(C1) * (k2 * k2 + k1 * k1) + 256) * sizeof(float);
  if (smem_staged <= 200 * 1024 && C1 % 2 == 0) {
    static size_t configured = 0;
    if (smem_staged > configured) {
      B2C_CUDA(cudaFuncSetAttribute(compose_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_staged)));
      configured = smem_staged;
    }
    B2C_CUDA(launch_k(compose_staged_kernel, dim3(C2), dim3(256), smem_staged, st, W1, b1, W2, b2, C1, C2, k1, s1, k2, kc, KP,
                      static_cast<__nv_bfloat16*>(Wc), Wc_f32, bc));
    B2C_CUDA(cudaGetLastError());
    return 0;
  }
  B2C_CUDA(launch_k(compose_w_kernel, dim3((C2 * KP * CS_SPLIT + 255) / 256), dim3(256), 0, st, W1, W2, C1, C2, k1, s1, k2, kc, KP,
                                                         static_cast<__nv_bfloat16*>(Wc), Wc_f32));
  B2C_CUDA(launch_k(compose_b_kernel, dim3(C2), dim3(32), 0, st, b1, W2, b2, C1, C2, k2, bc));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convsub_fwd(const float* x, const void* Wc, const float* bc, void* out, int out_f32, int B, int F,
                               int T, int C2, int kc, int stride, int KP, void* stream) {
  B2C_REQUIRE(x && Wc && bc && out, "b2c_convsub_fwd: null pointer");
  ConvSubParams p{};
  if (int rc = convsub_geometry(p, B, F, T, C2, kc, stride, KP, "b2c_convsub_fwd")) return rc;
  p.x = x; p.bc = bc; p.out = out; p.out_f32 = out_f32;
  CUtensorMap tm;
  if (int rc = get_tensor_map_2d(&tm, Wc, 2, KP, C2, static_cast<uint64_t>(KP) * 2, 64, C2, true)) return rc;
  const uint32_t smem = (KP / 64) * (16384 + C2 * 128) + 64 + 1024;
  static uint32_t configured = 0;
  if (smem > configured) {
    B2C_CUDA(cudaFuncSetAttribute(convsub_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(p.tiles_t, p.Fo, B);
  B2C_CUDA(launch_k(convsub_fwd_kernel, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), tm, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convsub_wgrad(const float* x, const void* dout, float* dWc, int B, int F, int T, int C2, int kc,
                                 int stride, int KP, void* stream) {
  B2C_REQUIRE(x && dout && dWc, "b2c_convsub_wgrad: null pointer");
  ConvSubParams p{};
  if (int rc = convsub_geometry(p, B, F, T, C2, kc, stride, KP, "b2c_convsub_wgrad")) return rc;
  B2C_REQUIRE(C2 % 128 == 0, "b2c_convsub_wgrad: out channels must be a multiple of 128");
  p.x = x; p.dWc = dWc;
  CUtensorMap tm;   // dout viewed as [B*To rows][Fo*C2] (token-major); MN-major A tiles: 64 channels x 128 frames
  if (int rc = get_tensor_map_2d(&tm, dout, 2, static_cast<uint64_t>(p.Fo) * C2, static_cast<uint64_t>(B) * p.To,
                                 static_cast<uint64_t>(p.Fo) * C2 * 2, 64, 128, true))
    return rc;
  const uint32_t smem = (KP / 64) * 16384 + 2 * 16384 + 64 + 1024;
  static uint32_t configured = 0;
  if (smem > configured) {
    B2C_CUDA(cudaFuncSetAttribute(convsub_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  int gx = num_sms() * 2 / (C2 / 128);
  if (gx > p.total_tiles) gx = p.total_tiles;
  dim3 grid(gx, C2 / 128);
  B2C_CUDA(launch_k(convsub_wgrad_kernel, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), tm, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_convsub_compose_bwd(const float* dWc, const float* dbc, const float* W1, const float* b1,
                                       const float* W2, int C1, int C2, int k1, int s1, int k2, int KP, float* dW1,
                                       float* db1, float* dW2, float* db2, void* stream) {
  B2C_REQUIRE(dWc && dbc && W1 && b1 && W2 && dW1 && db1 && dW2 && db2, "b2c_convsub_compose_bwd: null pointer");
  const int kc = k1 + (k2 - 1) * s1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n1 = C1 * (k1 * k1 + 1);
  const size_t smem_w1 = static_cast<size_t>(CB_OSLICE) * (kc * kc + 1) * sizeof(float);
  if (smem_w1 <= 48 * 1024)
    B2C_CUDA(launch_k(compose_bwd_w1_staged_kernel, dim3((n1 + 255) / 256, (C2 + CB_OSLICE - 1) / CB_OSLICE), dim3(256), smem_w1, st,
                      dWc, dbc, W2, C1, C2, k1, s1, k2, kc, KP, dW1, db1));
  else
    B2C_CUDA(launch_k(compose_bwd_w1_kernel, dim3((n1 * CS_SPLIT + 255) / 256), dim3(256), 0, st, dWc, dbc, W2, C1, C2, k1, s1, k2, kc, KP, dW1, db1));
  const int64_t n2 = static_cast<int64_t>(C2) * C1 * k2 * k2;
  B2C_CUDA(launch_k(compose_bwd_w2_kernel, dim3(static_cast<unsigned>((n2 + 127) / 128)), dim3(128), 0, st, dWc, dbc, W1, b1, C1, C2, k1, s1, k2,
                                                                                  kc, KP, dW2, db2));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
