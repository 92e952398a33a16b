// Fused macaron feed-forward half-step for INFERENCE (sm_100a, tcgen05 / TMEM / TMA):
//
//     out = x + alpha * ( W2 SiLU( W1 LayerNorm(x) + b1 ) + b2 )          alpha = 0.5
//
// i.e. torchaudio `_FeedForwardModule.forward` (conformer.py:102-109,118-119: LayerNorm, Linear, SiLU, Dropout, Linear,
// Dropout -- the dropouts are identities in eval mode) together with the `0.5 * x + residual` of `ConformerLayer.forward`
// (conformer.py:185-187,207-209), as ONE kernel.  Un-fused, the same half-step is a LayerNorm kernel and two GEMMs that
// write the [M, f] hidden activation to HBM and read it back (2 * M * f * 2 bytes: 98 MB per half-step at Conformer-M,
// 64 x 15 s) plus the bf16 LayerNorm output; SURVEY.md H1 predicted these skinny-K GEMMs stay HBM-bound unless the hidden
// activation stays on chip.  Here it never leaves the SM:
//
//   per 128-row tile of tokens (persistent CTAs, 576 threads):
//     16 compute warps   LayerNorm of the tile's rows in registers (warp per row) -> bf16 A operand in shared memory
//                        (K-major, 128B swizzle: exactly what TMA would have delivered)
//     warp 0 (1 lane)    TMA producer: streams W1 / W2 tiles ([128 rows][64 k], 16 KB) through an 8-stage ring, in the order the
//                        MMA warp consumes them
//     warp 1 (1 lane)    MMA issuer, software-pipelined over hidden chunks of 128 columns:
//                          H_j   = A  W1[j]^T   (M128 x N128, K = d)      -> TMEM acc1[j & 1]      (double-buffered)
//                          Y    += Hs_{j-1} W2[:, j-1]^T (M128 x N = d, K = 128)  -> TMEM acc2      (accumulates over j)
//     16 compute warps   mid-epilogue of chunk j: tcgen05.ld acc1 -> + b1 -> SiLU -> bf16 -> shared memory Hs[j & 1] in the
//                        UMMA operand layout (so it is the A operand of the second product), overlapping MMA of chunk j+1
//     16 compute warps   final epilogue: tcgen05.ld acc2 -> (+ b2) * alpha + x -> fp32 out
//   TMEM: 2 x 128 (acc1) + 256 (acc2) = 512 columns.  Shared memory (d = 256): A 64 KB + Hs 32 KB + ring 8 x 16 KB.
// Limits: d <= 256, d % 16 == 0 (Conformer-S 144, Conformer-M 256; Conformer-L, d = 512, keeps the un-fused path: its second
// accumulator alone would need all 512 TMEM columns), f % 8 == 0.
#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

constexpr int FF_BM = 128;          // token rows per tile
constexpr int FF_HC = 128;          // hidden columns per chunk
constexpr int FF_STAGES = 8;
constexpr int FF_CWARPS = 16;       // compute warps
constexpr int FF_THREADS = 64 + FF_CWARPS * 32;
constexpr uint32_t FF_STAGE_BYTES = 128 * 64 * 2;                 // every weight tile is [128 rows][64 k] bf16 = 16 KB

struct FfnParams {
  const float* x;
  float* out;
  const float* ln_w;
  const float* ln_b;
  const float* b1;
  const float* b2;
  int64_t M;
  int d, f, kb1, nj, m_tiles;     // kb1 = ceil(d / 64) k-blocks of the first product, nj = ceil(f / 128) hidden chunks
  float alpha, eps;
};

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1,
                                                                  const __grid_constant__ CUtensorMap tmW2,
                                                                  const FfnParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t A_BYTES = static_cast<uint32_t>(p.kb1) * 16384u;
  const uint32_t OFF_A = 0, OFF_HS = A_BYTES, OFF_W = OFF_HS + 32768u, OFF_BAR = OFF_W + FF_STAGES * FF_STAGE_BYTES;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_empty = w_full + FF_STAGES;
  uint64_t* a_full = w_empty + FF_STAGES;
  uint64_t* a_free = a_full + 1;
  uint64_t* acc1_full = a_free + 1;        // [2]
  uint64_t* acc1_free = acc1_full + 2;     // [2]
  uint64_t* hs_full = acc1_free + 2;       // Hs (single buffer) written by the mid-epilogue
  uint64_t* hs_free = hs_full + 1;         // second product of the previous chunk has consumed Hs
  uint64_t* acc2_full = hs_free + 1;
  uint64_t* acc2_free = acc2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < FF_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_free[s], FF_CWARPS);
    }
    mbar_init(hs_full, FF_CWARPS);
    mbar_init(hs_free, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, FF_CWARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_ACC2 = 256;
  pdl_wait();

  const int NJ = p.nj, KB1 = p.kb1;
  const int NH = (p.d + 127) / 128;           // 128-row halves of W2 (output columns of the second product)
  const int n_my_tiles = (p.m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ================================ TMA producer: W1 / W2 tiles in consumption order ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto push_w1 = [&](int j) {
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&w_empty[s], ph ^ 1u);
          mbar_expect_tx(&w_full[s], FF_HC * 64 * 2);
          tma_load_2d(smem + OFF_W + s * FF_STAGE_BYTES, &tmW1, &w_full[s], kb * 64, j * FF_HC);
          if (++s == FF_STAGES) { s = 0; ph ^= 1u; }
        }
      };
      auto push_w2 = [&](int j) {          // W2[:, hidden chunk j] as (k-block, 128-row half) tiles
        for (int kb = 0; kb < 2; ++kb)
          for (int hf = 0; hf < NH; ++hf) {
            mbar_wait(&w_empty[s], ph ^ 1u);
            mbar_expect_tx(&w_full[s], 128 * 64 * 2);
            tma_load_2d(smem + OFF_W + s * FF_STAGE_BYTES, &tmW2, &w_full[s], j * FF_HC + kb * 64, hf * 128);
            if (++s == FF_STAGES) { s = 0; ph ^= 1u; }
          }
      };
      for (int it = 0; it < n_my_tiles; ++it)
        for (int t = 0; t < NJ + 2; ++t) {      // the MMA warp's order: first product two chunks ahead of the second
          if (t < NJ) push_w1(t);
          if (t >= 2) push_w2(t - 2);
        }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_bf16(128, FF_HC, false, false);
      int s = 0;
      uint32_t ph = 0;
      uint32_t g0 = 0;                      // global index of this tile's first hidden chunk (slot = g & 1, use = g >> 1)
      auto mma1 = [&](uint32_t g, bool last) {        // H = A W1[chunk]^T  -> acc1[g & 1]
        const uint32_t as = g & 1u;
        mbar_wait(&acc1_free[as], ((g >> 1) & 1u) ^ 1u);       // mid-epilogue of chunk g-2 has read this accumulator
        tc_fence_after();
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&w_full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = base + OFF_A + kb * 16384u;
          const uint32_t b_addr = base + OFF_W + s * FF_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + as * FF_HC, umma_desc_sw128(a_addr + k * 32, 16, 1024),
                      umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&w_empty[s]);
          if (++s == FF_STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc1_full[as]);
        if (last) tc_commit(a_free);                  // all first-product MMAs of this tile issued: A is free once they retire
      };
      auto mma2 = [&](uint32_t g, bool first) {       // Y += Hs[g & 1] W2[:, chunk]^T, 128 output columns at a time
        mbar_wait(hs_full, g & 1u);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb)
          for (int hf = 0; hf < NH; ++hf) {
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            const uint32_t a_addr = base + OFF_HS + kb * 16384u;
            const uint32_t b_addr = base + OFF_W + s * FF_STAGE_BYTES;
            const int ncols = min(128, p.d - hf * 128);
            const uint32_t idesc2 = umma_idesc_bf16(128, ncols, false, false);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + TM_ACC2 + hf * 128, umma_desc_sw128(a_addr + k * 32, 16, 1024),
                        umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc2, (first && kb == 0 && k == 0) ? 0u : 1u);
            tc_commit(&w_empty[s]);
            if (++s == FF_STAGES) { s = 0; ph ^= 1u; }
          }
        tc_commit(hs_free);                 // the mid-epilogue may overwrite Hs once these MMAs retire
      };
      for (int it = 0; it < n_my_tiles; ++it) {
        mbar_wait(a_full, it & 1);           // LayerNorm output of this tile is in shared memory
        tc_fence_after();
        if (it > 0) {                        // the previous tile's final epilogue has drained acc2
          mbar_wait(acc2_free, (it - 1) & 1);
          tc_fence_after();
        }
        // software pipeline: the first product runs TWO chunks ahead of the second, so the mid-epilogue of a chunk (TMEM ->
        // SiLU -> shared memory) has two first-product main loops to hide under
        for (int t = 0; t < NJ + 2; ++t) {
          if (t < NJ) mma1(g0 + t, t == NJ - 1);
          if (t >= 2) mma2(g0 + t - 2, t == 2);
        }
        tc_commit(acc2_full);
        g0 += NJ;
      }
    }
  } else {
    // ================================ 16 compute warps ================================
    const int cw = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int part = cw >> 2;                // column slice
    const int d = p.d;
    uint32_t g = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      const int m_blk = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int64_t row0 = static_cast<int64_t>(m_blk) * FF_BM;
      // ---- LayerNorm prologue: warp per row, 8 rows per warp; lane holds columns lane*8 .. +8 ----
      if (it > 0) mbar_wait(a_free, (it - 1) & 1);        // previous tile's first-product MMAs have retired
      {
        const int c0 = lane * 8;
        float gw[8], gb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          gw[e] = c0 + e < d ? __ldg(p.ln_w + c0 + e) : 0.f;
          gb[e] = c0 + e < d ? __ldg(p.ln_b + c0 + e) : 0.f;
        }
        const float inv_d = 1.f / static_cast<float>(d);
        // all 8 rows of this warp are requested before the first is reduced: one memory round trip instead of eight
        float4 xa[8], xb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = min(row0 + cw * 8 + i, p.M - 1);                 // clamped: loads stay unconditional
          const int cl = c0 < d ? c0 : 0;
          xa[i] = *reinterpret_cast<const float4*>(p.x + row * d + cl);
          xb[i] = *reinterpret_cast<const float4*>(p.x + row * d + cl + 4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = cw * 8 + i;
          const int64_t row = row0 + r;
          const bool live = row < p.M && c0 < d;        // d % 8 == 0: a lane's 8 columns are all in or all out
          float v[8] = {xa[i].x, xa[i].y, xa[i].z, xa[i].w, xb[i].x, xb[i].y, xb[i].z, xb[i].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = live ? v[e] : 0.f;
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) sum += v[e];
          const float mean = warp_sum(sum) * inv_d;
          float sq = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dv = live ? v[e] - mean : 0.f;
            sq = fmaf(dv, dv, sq);
          }
          const float rstd = rsqrtf(warp_sum(sq) * inv_d + p.eps);
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = live ? fmaf((v[e] - mean) * rstd, gw[e], gb[e]) : 0.f;
          if (c0 < KB1 * 64) {                   // columns in [d, KB1*64) are zero padding of the last k-block
            const uint32_t off = static_cast<uint32_t>((c0 >> 6) * 16384 + r * 128 + (((((c0 & 63) >> 3)) ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(smem + OFF_A + off) =
                make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          }
        }
      }
      fence_proxy_async();                         // generic-proxy stores -> visible to the UMMA async proxy
      named_barrier_sync(1, FF_CWARPS * 32);
      if (cw == 0 && lane == 0) mbar_arrive(a_full);

      // ---- mid-epilogue per hidden chunk: acc1 -> + b1 -> SiLU -> bf16 -> Hs (A operand of the second product) ----
      for (int j = 0; j < NJ; ++j, ++g) {
        const uint32_t as = g & 1u;
        mbar_wait(&acc1_full[as], (g >> 1) & 1u);
        tc_fence_after();
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tmem_base + as * FF_HC + (static_cast<uint32_t>(q * 32) << 16) + part * 32, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_free[as]);
        const int hcol0 = j * FF_HC + part * 32;             // first hidden column of this thread's 32
        const int r = q * 32 + lane;
        uint4 hv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = hcol0 + c * 8 + e;
            const float z = __uint_as_float(v[c * 8 + e]) + (col < p.f ? __ldg(p.b1 + col) : 0.f);
            h[e] = col < p.f ? silu_f(z) : 0.f;
          }
          hv[c] = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
        }
        mbar_wait(hs_free, (g & 1u) ^ 1u);                   // second product of chunk g-1 has consumed Hs
        uint8_t* hs = smem + OFF_HS + (part >> 1) * 16384u + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = (part & 1) * 4 + c;
          *reinterpret_cast<uint4*>(hs + ((chunk ^ (r & 7)) << 4)) = hv[c];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(hs_full);
      }

      // ---- final epilogue: acc2 -> (+ b2) * alpha + x -> out (thread = row, its slice of the columns) ----
      // the residual row pieces are requested BEFORE the accumulator is awaited: their latency hides under the last MMAs
      const int span = ((d + 63) / 64) * 16;               // columns per warp slice (4 slices), multiple of 16, <= 64
      const int64_t row = row0 + q * 32 + lane;
      const int64_t rowc = min(row, p.M - 1);
      float4 xr[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = part * span + i * 4;
        if (i * 4 < span) xr[i] = ldg_stream_f4(p.x + rowc * d + (col < d ? col : 0));
      }
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        if (cc < span) {
          const int col0 = part * span + cc;
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tmem_base + TM_ACC2 + (static_cast<uint32_t>(q * 32) << 16) + col0, v);
          tmem_wait_ld();
          if (row < p.M && col0 < d) {
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) {
              if (col0 + c4 < d) {                            // d % 4 == 0
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + c4));
                const float4 x4 = xr[(cc + c4) >> 2];
                float4 o;
                o.x = fmaf(p.alpha, __uint_as_float(v[c4]) + bb.x, x4.x);
                o.y = fmaf(p.alpha, __uint_as_float(v[c4 + 1]) + bb.y, x4.y);
                o.z = fmaf(p.alpha, __uint_as_float(v[c4 + 2]) + bb.z, x4.z);
                o.w = fmaf(p.alpha, __uint_as_float(v[c4 + 3]) + bb.w, x4.w);
                *reinterpret_cast<float4*>(p.out + row * d + col0 + c4) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_ffn_fused_eval(const float* x, float* out, const float* ln_w, const float* ln_b, const void* w1,
                                  const float* b1, const void* w2, const float* b2, int64_t M, int d, int f, float alpha,
                                  float eps, void* stream) {
  B2C_REQUIRE(x && out && ln_w && ln_b && w1 && b1 && w2 && b2, "b2c_ffn_fused_eval: null pointer");
  B2C_REQUIRE(M > 0 && d > 0 && f > 0, "b2c_ffn_fused_eval: empty problem");
  B2C_REQUIRE(d % 16 == 0 && d <= 256, "b2c_ffn_fused_eval: d=%d must be a multiple of 16 and <= 256 (use the un-fused path otherwise)", d);
  B2C_REQUIRE(f % 8 == 0, "b2c_ffn_fused_eval: f=%d must be a multiple of 8", f);
  B2C_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0,
              "b2c_ffn_fused_eval: x / out / b2 must be 16-byte aligned");
  FfnParams p;
  p.x = x; p.out = out; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2;
  p.M = M; p.d = d; p.f = f; p.alpha = alpha; p.eps = eps;
  p.kb1 = (d + 63) / 64;
  p.nj = (f + FF_HC - 1) / FF_HC;
  p.m_tiles = static_cast<int>((M + FF_BM - 1) / FF_BM);
  CUtensorMap t1, t2;
  if (int rc = get_tensor_map_2d(&t1, w1, 2, static_cast<uint64_t>(d), static_cast<uint64_t>(f), static_cast<uint64_t>(d) * 2, 64, FF_HC, true)) return rc;
  if (int rc = get_tensor_map_2d(&t2, w2, 2, static_cast<uint64_t>(f), static_cast<uint64_t>(d), static_cast<uint64_t>(f) * 2, 64, 128, true)) return rc;
  const uint32_t smem = static_cast<uint32_t>(p.kb1) * 16384u + 32768u + FF_STAGES * FF_STAGE_BYTES + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 32768 + FF_STAGES * FF_STAGE_BYTES + 256 + 1024));
    configured = true;
  }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  B2C_CUDA(launch_k(ffn_fused_kernel, dim3(grid), dim3(FF_THREADS), smem, static_cast<cudaStream_t>(stream), t1, t2, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
