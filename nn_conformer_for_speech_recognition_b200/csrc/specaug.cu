// SpecAugment (/root/reference/lib/standard/asrnn.py:91-192): time-warp gather + frequency / time masks, one pass.
//
// The reference builds the warp table with a Python loop per frame (a torch.div(...).item() each), gathers with numpy
// per (utterance, mel bin) on the CPU and re-uploads (asrnn.py:102-124) -- 212 ms for a 4x80x1000 batch.  Here the
// draws stay on the host (they come from CPython's global `random` stream and must stay bit-exact) and the data
// never leaves the GPU: thread = (utterance, frame) computes its source frame once -- the same IEEE double
// divide/multiply/truncate and int64 floor division Python performs -- then walks the mel bins with coalesced
// loads/stores.  HBM-bound: algorithmic bytes = 2 * B*F*T * 4.
#include "common.cuh"
#include "b200conformer.h"

namespace b2c {

__device__ __forceinline__ long long floordiv_ll(long long a, long long b) {
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

// Wt of asrnn.py:108-115 for frame t of an utterance with length tau
__device__ __forceinline__ long long warp_src(int t, int w, int w0, int tau) {
  if (t >= tau) return t;                                              // asrnn.py:115 identity tail
  if (t <= w0) {
    const double r = static_cast<double>(w0 + w) / static_cast<double>(w0);   // Python: (w0+w)/w0, true division
    return static_cast<long long>(r * static_cast<double>(t));         // int(): truncation
  }
  const long long num = static_cast<long long>(tau - 1 - w0 - w) * t + static_cast<long long>(tau - 1) * w;
  return floordiv_ll(num, static_cast<long long>(tau - 1 - w0));       // torch.div(..., rounding_mode='floor')
}

__global__ void __launch_bounds__(256) specaug_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                     const int32_t* __restrict__ tau, const int32_t* __restrict__ warp_w,
                                                     const int32_t* __restrict__ warp_w0, int n_warp,
                                                     const int32_t* __restrict__ fmask, int n_fmask,
                                                     const int32_t* __restrict__ tmask, int n_tmask, int apply_masks,
                                                     float mask_value, int B, int F, int T, int f_per_block) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const int f_lo = blockIdx.z * f_per_block, f_hi = min(F, f_lo + f_per_block);
  if (t >= T) return;
  const int len = tau[b];
  long long src = t;
  for (int pass = n_warp - 1; pass >= 0; --pass) {                      // x_n[t] = x_{n-1}[W_n[t]]  => compose backwards
    src = warp_src(static_cast<int>(src), warp_w[pass * B + b], warp_w0[pass * B + b], len);
    src = src < 0 ? 0 : (src >= T ? T - 1 : src);
  }
  bool t_masked = false;
  if (apply_masks) {
    for (int n = 0; n < n_tmask; ++n) {
      const int t0 = tmask[(n * B + b) * 2], tl = tmask[(n * B + b) * 2 + 1];
      t_masked |= (t >= t0 && t < t0 + tl);
    }
  }
  const float* xb = x + static_cast<int64_t>(b) * F * T;
  float* ob = out + static_cast<int64_t>(b) * F * T;
#pragma unroll 4
  for (int f = f_lo; f < f_hi; ++f) {
    bool masked = t_masked;
    if (apply_masks) {
      for (int n = 0; n < n_fmask; ++n) masked |= (f >= fmask[2 * n] && f < fmask[2 * n] + fmask[2 * n + 1]);
    }
    ob[static_cast<int64_t>(f) * T + t] = masked ? mask_value : __ldg(xb + static_cast<int64_t>(f) * T + src);
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_specaug_apply(const float* x, float* out, const int32_t* tau, const int32_t* warp_w,
                                 const int32_t* warp_w0, int n_warp, const int32_t* fmask, int n_fmask,
                                 const int32_t* tmask, int n_tmask, int apply_masks, float mask_value, int B, int F,
                                 int T, void* stream) {
  B2C_REQUIRE(x && out && tau, "b2c_specaug_apply: null pointer");
  B2C_REQUIRE(x != out, "b2c_specaug_apply: in-place operation is not supported (gather)");
  B2C_REQUIRE(n_warp == 0 || (warp_w && warp_w0), "b2c_specaug_apply: warp parameters missing");
  B2C_REQUIRE(n_fmask == 0 || fmask, "b2c_specaug_apply: fmask missing");
  B2C_REQUIRE(n_tmask == 0 || tmask, "b2c_specaug_apply: tmask missing");
  if (B <= 0 || F <= 0 || T <= 0) return 0;
  const int fz = F >= 32 ? 4 : 1;
  const int f_per_block = (F + fz - 1) / fz;
  dim3 grid((T + 255) / 256, B, (F + f_per_block - 1) / f_per_block);
  B2C_CUDA(launch_k(specaug_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, tau, warp_w, warp_w0, n_warp, fmask,
                                                                      n_fmask, tmask, n_tmask, apply_masks, mask_value,
                                                                      B, F, T, f_per_block));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// CPython-compatible MT19937 (host).  state[0..623] = mt, state[624] = index.
// ------------------------------------------------------------------------------------------------------------
static void mt_init_genrand(uint32_t* mt, uint32_t s) {
  mt[0] = s;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
  mt[624] = 624;
}

extern "C" void b2c_mt19937_seed_host(uint32_t* st, uint64_t seed) {
  uint32_t key[2] = {static_cast<uint32_t>(seed & 0xFFFFFFFFu), static_cast<uint32_t>(seed >> 32)};
  const int klen = key[1] ? 2 : 1;                    // random.seed(int): 32-bit limbs of |seed|, little endian
  mt_init_genrand(st, 19650218u);
  int i = 1, j = 0;
  for (int k = (624 > klen ? 624 : klen); k; --k) {
    st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1664525u)) + key[j] + static_cast<uint32_t>(j);
    ++i; ++j;
    if (i >= 624) { st[0] = st[623]; i = 1; }
    if (j >= klen) j = 0;
  }
  for (int k = 623; k; --k) {
    st[i] = (st[i] ^ ((st[i - 1] ^ (st[i - 1] >> 30)) * 1566083941u)) - static_cast<uint32_t>(i);
    ++i;
    if (i >= 624) { st[0] = st[623]; i = 1; }
  }
  st[0] = 0x80000000u;
  st[624] = 624;
}

static uint32_t mt_u32(uint32_t* st) {
  if (st[624] >= 624) {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (st[k] & 0x80000000u) | (st[(k + 1) % 624] & 0x7FFFFFFFu);
      st[k] = st[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    }
    st[624] = 0;
  }
  uint32_t y = st[st[624]++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

// random.randint(a, b) for b - a + 1 <= 2^32: a + _randbelow(n), n.bit_length() bits per try, reject r >= n
extern "C" int64_t b2c_mt19937_randint_host(uint32_t* st, int64_t a, int64_t b) {
  const uint64_t n = static_cast<uint64_t>(b - a + 1);
  if (b < a || n > 0x100000000ull) return INT64_MIN;  // Python raises ValueError on an empty range
  int k = 0;
  for (uint64_t m = n; m; m >>= 1) ++k;
  uint64_t r;
  do {
    r = k == 0 ? 0 : (mt_u32(st) >> (32 - (k > 32 ? 32 : k)));
    if (k > 32) r = (r << 1);                        // n == 2^32: unreachable for SpecAugment ranges; kept total
  } while (r >= n);
  return a + static_cast<int64_t>(r);
}
