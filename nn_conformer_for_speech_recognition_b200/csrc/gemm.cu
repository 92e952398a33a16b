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
// (torch/nn/functional.py:5798,6690), with the bias / SiLU / 0.5*residual elementwise kernels fused as epilogue.
//
// Structure: one 128 x BN output tile per CTA, 192 threads:
//   warp 0    TMA producer (one elected lane)  -> 128B-swizzled smem ring, mbarrier full/empty
//   warp 1    TMEM allocator + MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16
//   warps 2-5 epilogue: tcgen05.ld 32x32b (thread = accumulator row) -> bias/act/residual/dropout -> global
// 2-3 CTAs are co-resident per SM (TMEM BN columns each), so one CTA's epilogue overlaps another's main loop.
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmParams {
  int64_t M, N, K;
  int kblocks_per_split;
  b2c_epilogue ep;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 3 : 4;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* accum_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int total_kb = static_cast<int>((p.K + BK - 1) / BK);
  const int kb0 = blockIdx.z * p.kblocks_per_split;
  const int kb1 = min(total_kb, kb0 + p.kblocks_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        if constexpr (!A_MN) {
          tma_load_2d(sA, &tmA, &full_bar[s], kb * BK, m_blk * BM);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(sA + c * 8192, &tmA, &full_bar[s], m_blk * BM + c * 64, kb * BK);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sB, &tmB, &full_bar[s], kb * BK, n_blk * BN);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sB + c * 8192, &tmB, &full_bar[s], n_blk * BN + c * 64, kb * BK);
        }
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = base + s * Cfg::STAGE_BYTES;
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = A_MN ? umma_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                   : umma_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                   : umma_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
      }
      tc_commit(accum_bar);
    }
  } else {
    // ------------------------------ epilogue: thread <-> accumulator row ------------------------------
    const b2c_epilogue& ep = p.ep;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int64_t row = static_cast<int64_t>(m_blk) * BM + q * 32 + lane;
    const bool row_ok = row < p.M;
    const float* bias = blockIdx.z == 0 ? static_cast<const float*>(ep.bias) : nullptr;   // once per split-K
    const uint32_t drop_thresh = ep.drop_p > 0.f ? static_cast<uint32_t>(ep.drop_p * 4294967296.0) : 0u;
    const float inv_keep = ep.drop_p > 0.f ? 1.f / (1.f - ep.drop_p) : 1.f;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      __syncwarp();                            // tcgen05.ld is .sync.aligned: reconverge after masked stores
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      tmem_wait_ld();
      const int64_t col0 = static_cast<int64_t>(n_blk) * BN + c;
      if (!row_ok || col0 >= p.N) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {            // 4 groups of 8 columns
        const int64_t col = col0 + g * 8;
        if (col >= p.N) break;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
        if (bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
          f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if (ep.out2) {                         // pre-activation copy (saved for the backward pass)
          uint4 u = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out2) + row * ep.ld_out2 + col) = u;
        }
        if (ep.act == B2C_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
        } else if (ep.act == B2C_ACT_SILU_BWD) {   // f = dL/dh ; aux = pre-activation u ; out = f * silu'(u)
          const uint4 a = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(ep.aux) + row * ep.ld_aux + col);
          const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
          const float u[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] *= silu_grad_f(u[j]);
        }
        if (drop_thresh) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            f[j] *= drop_scale(ep.drop_seed, static_cast<uint64_t>(row) * p.N + col + j, drop_thresh, inv_keep);
        }
        if (ep.alpha != 1.f) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] *= ep.alpha;
        }
        if (ep.out_dtype == B2C_BF16) {
          uint4 u = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + row * ep.ld_out + col) = u;
        } else {
          float* o = static_cast<float*>(ep.out) + row * ep.ld_out + col;
          if (ep.accumulate) {                 // split-K / gradient accumulation: fp32 reductions in L2
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(o + j, f[j]);
          } else {
            if (ep.resid) {
              const float* r = static_cast<const float*>(ep.resid) + row * ep.ld_out + col;
              const float4 r0 = *reinterpret_cast<const float4*>(r);
              const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
              f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
              f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
            }
            *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    B2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>((p.M + BM - 1) / BM), static_cast<unsigned>((p.N + BN - 1) / BN), splits);
  kern<<<grid, 192, Cfg::SMEM, st>>>(ta, tb, p);
  B2C_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                             int64_t N, int64_t K, int split_k, int block_n, const b2c_epilogue* ep, void* stream) {
  B2C_REQUIRE(A && B && ep && ep->out, "b2c_gemm_bf16: null pointer");
  B2C_REQUIRE(M > 0 && N > 0 && K > 0, "b2c_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N,
              (long long)K);
  B2C_REQUIRE(N % 8 == 0, "b2c_gemm_bf16: N=%lld must be a multiple of 8", (long long)N);
  B2C_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "b2c_gemm_bf16: leading dims must be multiples of 8 elements (16 B)");
  B2C_REQUIRE(ep->ld_out % 4 == 0, "b2c_gemm_bf16: ld_out must be a multiple of 4");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0,
              "b2c_gemm_bf16: pointers must be 16-byte aligned");
  B2C_REQUIRE(!(ep->accumulate && ep->out_dtype != B2C_F32), "b2c_gemm_bf16: accumulate needs fp32 output");
  B2C_REQUIRE(!(ep->act == B2C_ACT_SILU_BWD && !ep->aux), "b2c_gemm_bf16: SILU_BWD needs aux");
  int bn = block_n;
  if (bn == 0) bn = (N >= 128) ? 128 : 64;
  B2C_REQUIRE(bn == 64 || bn == 128 || bn == 256, "b2c_gemm_bf16: block_n must be 64, 128 or 256");
  const int total_kb = static_cast<int>((K + BK - 1) / BK);
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > total_kb) splits = total_kb;
  B2C_REQUIRE(splits == 1 || ep->accumulate, "b2c_gemm_bf16: split_k > 1 needs accumulate=1 (fp32 atomics)");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.kblocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.ep = *ep;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = get_tensor_map_2d(&ta, A, 2, K, M, lda * 2, BK, BM, true);
  else       rc = get_tensor_map_2d(&ta, A, 2, M, K, lda * 2, 64, BK, true);
  if (rc) return rc;
  if (!b_mn) rc = get_tensor_map_2d(&tb, B, 2, K, N, ldb * 2, BK, bn, true);
  else       rc = get_tensor_map_2d(&tb, B, 2, N, K, ldb * 2, 64, BK, true);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2C_DISPATCH(BN_)                                                                 \
  if (bn == BN_) {                                                                        \
    if (!a_mn && !b_mn) return launch_gemm<BN_, false, false>(ta, tb, p, splits, st);     \
    if (!a_mn && b_mn) return launch_gemm<BN_, false, true>(ta, tb, p, splits, st);       \
    if (a_mn && b_mn) return launch_gemm<BN_, true, true>(ta, tb, p, splits, st);         \
    return launch_gemm<BN_, true, false>(ta, tb, p, splits, st);                          \
  }
  B2C_DISPATCH(64)
  B2C_DISPATCH(128)
  B2C_DISPATCH(256)
#undef B2C_DISPATCH
  return -1;
}
