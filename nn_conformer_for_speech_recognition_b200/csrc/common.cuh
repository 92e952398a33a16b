// Shared device/host helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// UMMA descriptors, counter-based dropout hash, error plumbing for the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2c {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);      // records message, returns (int)e
#define B2C_CUDA(call)                                              \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::b2c::cuda_fail(e__, #call);    \
  } while (0)
#define B2C_REQUIRE(cond, ...)                                      \
  do {                                                              \
    if (!(cond)) { ::b2c::set_error(__VA_ARGS__); return -1; }      \
  } while (0)

// Encodes (or fetches from a per-process cache) a 2-D bf16/fp32 TMA tensor map.
//   inner/outer: extents in elements, ld_bytes: byte pitch of the outer dim, box: {inner, outer} elements.
//   swizzle128: CU_TENSOR_MAP_SWIZZLE_128B when true, none otherwise.
int get_tensor_map_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128);

// same with an explicit swizzle mode: 0 = none, 1 = 128 B, 2 = 64 B
int get_tensor_map_2d_sw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                         uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle);
int get_tensor_map_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, bool swizzle128);

int num_sms();
bool pdl_enabled();       // host.cpp: programmatic dependent launch on unless B2C_PDL=0

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The step is ~750 short kernels in one stream / CUDA graph; between two
// dependent kernels the GPU otherwise drains completely, then pays the launch latency and the next kernel's prologue
// (barrier init, TMEM allocation, descriptor prefetch) with every SM idle.  Every kernel of this library is launched with
// programmaticStreamSerializationAllowed: its CTAs become resident while the previous kernel's last wave still runs, execute
// their prologue, and block in pdl_wait() until the previous grid has COMPLETED and its memory is visible.
// Contract for every kernel launched through launch_k():
//   pdl_trigger() first (lets the NEXT kernel start being scheduled as soon as all CTAs of this one have started), and
//   pdl_wait() before the first global-memory access that is not a kernel parameter / tensor map.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// sigmoid through ONE MUFU.TANH: 0.5*tanh(0.5 z) + 0.5.  tanh.approx has ~2^-11 relative error, far below the bf16
// rounding (2^-9) every consumer of these values applies; it replaces an exp + full-precision divide (~15 instructions).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float z) { return fmaf(0.5f, tanh_approx(0.5f * z), 0.5f); }
__device__ __forceinline__ float silu_f(float z) { return z * sigmoid_f(z); }
// d/dz [z*sigmoid(z)] = s + z s (1 - s), s = (1 + t)/2, t = tanh(z/2)  =>  (1 + t + (z/2)(1 - t^2)) / 2: one MUFU, three FFMA
__device__ __forceinline__ float silu_grad_f(float z) {
  const float hz = 0.5f * z;
  const float t = tanh_approx(hz);
  return fmaf(0.5f, fmaf(hz, fmaf(-t, t, 1.f), t), 0.5f);
}
// packed fp32 pairs kept in 64-bit registers (FADD2 / FMUL2 / FFMA2: half the issue slots of the scalar forms); pk2 / upk2
// are register-pair renames, not instructions, when the halves already sit in an aligned pair (tcgen05.ld / LDS.128 results)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// bf16x2 word from a packed fp32 pair (low half = low lane)
__device__ __forceinline__ uint32_t cvt_bf16x2(uint64_t v) {
  float lo, hi;
  upk2(v, lo, hi);
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// 0xFFFF / 0x0000 per 16-bit half from bit 7 (low half) and bit 15 (high half) of w: PRMT with sign replication
__device__ __forceinline__ uint32_t halfmask_from_bits_7_15(uint32_t w) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(0x9988u));
  return d;
}
// packed fp32 FMA (FFMA2): d = a * b + c on two lanes of a 64-bit register pair -- half the issue slots of two FFMAs
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  ra = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  rb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  rc = (static_cast<uint64_t>(__float_as_uint(c.y)) << 32) | __float_as_uint(c.x);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return make_float2(__uint_as_float(static_cast<uint32_t>(rd)), __uint_as_float(static_cast<uint32_t>(rd >> 32)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Counter-based dropout.  One 32-bit hash serves the element PAIR (idx, idx+1), idx even: 16 bits per decision,
// keep when the 16-bit lane >= thresh16 = round(p * 65536).  Forward and backward regenerate identical masks from
// (seed, idx); index spaces are laid out with even row pitches so pairs never straddle rows.
// Mixing: per-tensor affine map of the pair index (odd multiplier and offset derived from the seed) followed by the
// "lowbias32" avalanche (2 multiplies, 3 xor-shifts) -- ~4 integer ops per element instead of a 64-bit splitmix.
__host__ __device__ __forceinline__ uint32_t drop_hash32(uint64_t seed, uint64_t pair_idx) {
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t x = static_cast<uint32_t>(pair_idx) * (k1 | 1u) + k0 + static_cast<uint32_t>(pair_idx >> 32) * 0x85EBCA6Bu;
  x ^= x >> 16; x *= 0x21F0AAADu;
  x ^= x >> 15; x *= 0x735A2D97u;
  x ^= x >> 15;
  return x;
}
struct DropCfg {
  uint64_t seed;             // well-mixed 64-bit key (host side runs the user seed through splitmix64)
  const uint64_t* epoch;     // device counter bumped once per training step (b2c_dropout_epoch_bump): lets a CUDA graph
                             // replay the same launches with fresh masks; forward and backward of a step see one value
  uint32_t thresh16;         // 0 = dropout off
  float inv_keep;
  float inv_keep8;           // 256 / (256 - (thresh16 >> 8)): scale of the 8-bit-per-decision variants
};
uint64_t* dropout_epoch_ptr();             // dropout.cu: the CURRENT DEVICE's step counter (lazily allocated, one per device)
const uint64_t* dropout_epoch_current();   // the word kernels launched by this thread read: the one installed with
                                           // b2c_dropout_epoch_use (a per-call snapshot), else the device counter
__host__ __device__ __forceinline__ uint64_t drop_mix_seed(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline DropCfg make_drop(float p, uint64_t seed) {
  DropCfg d;
  d.seed = drop_mix_seed(seed);
  d.epoch = p > 0.f ? dropout_epoch_current() : nullptr;
  d.thresh16 = p > 0.f ? static_cast<uint32_t>(p * 65536.0 + 0.5) : 0u;
  if (d.thresh16 > 65535u) d.thresh16 = 65535u;
  d.inv_keep = p > 0.f ? 65536.f / static_cast<float>(65536u - d.thresh16) : 1.f;
  d.inv_keep8 = 256.f / (256.f - static_cast<float>(d.thresh16 >> 8));
  return d;
}
// fold the step epoch into the key (call once per thread, before the first drop_scale2)
__device__ __forceinline__ DropCfg drop_resolve(DropCfg d) {
  if (d.thresh16 && d.epoch) d.seed += (*d.epoch) * 0x9E3779B97F4A7C15ull;
  return d;
}
// single-instruction 2^x (MUFU.EX2, flush-to-zero): exp2f() adds range fix-ups the softmax does not need
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// scales for elements idx (multiple of 4) .. idx+3 from ONE hash, 8 bits per decision (attention probabilities: the
// per-element budget there is a handful of instructions).  thresh8 = thresh16 >> 8; exact for p = 0.5, and
// inv_keep8 = 256 / (256 - thresh8) keeps the mask unbiased for any p.
__device__ __forceinline__ void drop_scale4(const DropCfg& d, uint64_t idx4, float inv_keep8, float* s) {
  const uint32_t h = drop_hash32(d.seed ^ 0x5851F42D4C957F2Dull, idx4 >> 2);
  const uint32_t t8 = d.thresh16 >> 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) s[j] = ((h >> (8 * j)) & 0xFFu) >= t8 ? inv_keep8 : 0.f;
}
// Same idea for the elementwise sites (GEMM epilogues, scale/cast passes): ONE hash per 4 consecutive elements
// (idx4 a multiple of 4), 8 bits per decision -- the epilogues are issue bound and the per-pair hash was ~5 integer
// instructions per element.  Forward and backward of a site call this same function, so masks always agree.
__device__ __forceinline__ void drop_scale4g(const DropCfg& d, uint64_t idx4, float& s0, float& s1, float& s2, float& s3) {
  const uint32_t h = drop_hash32(d.seed, idx4 >> 2);
  const uint32_t t8 = d.thresh16 >> 8;
  const float k = d.inv_keep8;
  s0 = (h & 0xFFu) >= t8 ? k : 0.f;
  s1 = ((h >> 8) & 0xFFu) >= t8 ? k : 0.f;
  s2 = ((h >> 16) & 0xFFu) >= t8 ? k : 0.f;
  s3 = (h >> 24) >= t8 ? k : 0.f;
}
// attention probabilities at p = 0.5 (the reference's dropout): ONE hash decides 32 consecutive keys of a query row, one
// bit each.  `row_lin` = (b*H + h)*T + q, `key32` = first key of the aligned 32-key group.  Forward, backward and the
// mask export (b2c_dropout_export_mask) all call this, so the masks agree.
__device__ __forceinline__ uint32_t drop_bits32(const DropCfg& d, uint64_t row_lin, int T, int key32) {
  return drop_hash32(d.seed ^ 0x94D049BB133111EBull, row_lin * static_cast<uint64_t>((T + 31) >> 5) + static_cast<uint64_t>(key32 >> 5));
}
// scales for elements idx (even) and idx+1
__device__ __forceinline__ void drop_scale2(const DropCfg& d, uint64_t idx_even, float& s0, float& s1) {
  const uint32_t h = drop_hash32(d.seed, idx_even >> 1);
  s0 = (h & 0xFFFFu) >= d.thresh16 ? d.inv_keep : 0.f;
  s1 = (h >> 16) >= d.thresh16 ? d.inv_keep : 0.f;
}

// streaming global loads that do not allocate in L1 (epilogue operands are read once; with ~200 KB of the SM's SRAM
// carved out as shared memory the remaining L1 holds only ~200 lines)
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  const uint4 v = ldg_stream_v4(p);
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps (-> launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// pull one box of a tensor into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
// 2-D tile load, coordinates {c_inner, c_outer}; completes `bar` with the box's bytes.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// multicast tile load: the box lands at the SAME smem offset in every CTA of the cluster whose bit is set in `mask`, and
// completes the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "h"(mask)
      : "memory");
}
// 3-D tile STORE shared -> global (bulk async group); rows / columns outside the tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 3-D tile REDUCE-ADD shared -> global: out[box] += smem[box], performed by the TMA unit at the L2 (fp32 per the tensor
// map); clipped like a store.  Lets several CTAs accumulate partial tiles into ONE buffer without a slab round trip.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the N most recently committed bulk-store groups of this thread have finished reading shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this smem offset in every CTA of the cluster selected by `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns (one fp32 column per register); thread i of the warp reads TMEM lane base+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 columns store (thread i writes TMEM lane base+i); pair with tmem_wait_st() before re-reading
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// cta_group::2 (CTA pair on one TPC): the leader CTA issues MMAs of M = 256 whose operands come from BOTH CTAs'
// shared memory (each holds its 128 rows of A and half of B's N rows) and whose accumulator rows split across both
// CTAs' TMEM.  Per output element only half the operand bytes cross the L2 -> SMEM fabric.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load issued by either CTA of a pair; completes the LEADER's mbarrier (cluster address `bar_cluster_addr`)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c_inner,
                                                int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retire) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bf16, SWIZZLE_128B tiles whose rows are 128 bytes)
//   K-major  : smem holds [rows][64 k] (128 B per row, 8-row groups of 1024 B): SBO = 1024 B, k-step = +32 B
//   MN-major : smem holds [k][64 mn]  (128 B per k-row), 64-wide MN chunks `lbo_bytes` apart:
//              SBO = 1024 B (8 k-rows), LBO = chunk pitch, k-step (16 k-rows) = +2048 B
// bit layout: addr>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 (SW128) [61,64)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor: f32 accumulate, bf16 A/B
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace b2c
