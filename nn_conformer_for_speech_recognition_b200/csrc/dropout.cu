// Dropout epoch plumbing and the mask-export debug entry.
//
// Every dropout site of the hot path (torchaudio conformer.py:74,106,108,152,153; asrnn.py:210) draws its mask as a pure
// function of (seed, element index, epoch word).  The epoch word is read by the KERNEL at run time, which is what lets a
// replayed CUDA graph draw fresh masks each step.  Two levels:
//   * one step counter per DEVICE (lazily allocated on the device that is current at first use);
//   * a per-CALL snapshot: a training forward copies ++counter into a word it owns (b2c_dropout_epoch_snapshot) and
//     installs that word for the launches of its forward AND its backward (b2c_dropout_epoch_use).  Two forwards
//     followed by two backwards (summed losses, teacher / student pairs, recomputation) therefore regenerate the masks
//     each forward really used -- with a single global word the earlier call's backward would see the later epoch.
// b2c_dropout_export_mask materialises the masks with the SAME device functions the kernels call: the parity tests
// replay them in an fp64 reference layer (tests/test_dropout_gpu.py).
#include <mutex>

#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

constexpr int MAX_DEV = 64;
static uint64_t* g_epoch[MAX_DEV] = {};
static std::mutex g_epoch_mu;
static thread_local const uint64_t* g_epoch_use = nullptr;

uint64_t* dropout_epoch_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
  std::lock_guard<std::mutex> g(g_epoch_mu);
  if (!g_epoch[dev]) {
    uint64_t* p = nullptr;
    if (cudaMalloc(&p, sizeof(uint64_t)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(uint64_t));
    g_epoch[dev] = p;
  }
  return g_epoch[dev];
}

const uint64_t* dropout_epoch_current() { return g_epoch_use ? g_epoch_use : dropout_epoch_ptr(); }

__global__ void epoch_bump_kernel(uint64_t* e, uint64_t* snap) {
  pdl_trigger();
  pdl_wait();
  const uint64_t v = *e + 1;
  *e = v;
  if (snap) *snap = v;
}
__global__ void epoch_set_kernel(uint64_t* e, uint64_t v) {
  pdl_trigger();
  pdl_wait(); *e = v; }

// kind 0: elementwise sites -- the scale drop_scale4g applies to element (r, c) of a [rows, cols] tensor
//         (GEMM residual epilogues, scale_cast, layernorm_bwd_cast; index = r*cols + c)
// kind 1: attention probabilities -- rows = B*H*T query rows (row_lin = (b*H + h)*T + q), cols = T keys
__global__ void export_mask_kernel(int kind, DropCfg drop, int64_t rows, int64_t cols, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  drop = drop_resolve(drop);
  const int64_t cp = (cols + 3) & ~static_cast<int64_t>(3);      // the attention kernels' even row pitch (Tpad)
  const int64_t n4 = rows * (cp >> 2);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / (cp >> 2), c = (i % (cp >> 2)) * 4;
    float s[4];
    if (kind == 0) {
      drop_scale4g(drop, static_cast<uint64_t>(r * cols + c), s[0], s[1], s[2], s[3]);
    } else if (drop.thresh16 == 32768u) {
      const uint32_t bits = drop_bits32(drop, static_cast<uint64_t>(r), static_cast<int>(cols), static_cast<int>(c) & ~31);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = (bits >> ((c & 31) + e)) & 1u ? 2.f : 0.f;
    } else {
      drop_scale4(drop, static_cast<uint64_t>(r) * static_cast<uint64_t>(cp) + static_cast<uint64_t>(c), drop.inv_keep8, s);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (c + e < cols) out[r * cols + c + e] = s[e];
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_dropout_epoch_bump(void* stream) {
  uint64_t* e = dropout_epoch_ptr();
  B2C_REQUIRE(e != nullptr, "b2c_dropout_epoch_bump: could not allocate the epoch word");
  B2C_CUDA(launch_k(epoch_bump_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), e, nullptr));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_dropout_epoch_snapshot(uint64_t* dst, void* stream) {
  uint64_t* e = dropout_epoch_ptr();
  B2C_REQUIRE(e != nullptr && dst != nullptr, "b2c_dropout_epoch_snapshot: null pointer");
  B2C_CUDA(launch_k(epoch_bump_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), e, dst));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_dropout_epoch_set(uint64_t value, void* stream) {
  uint64_t* e = dropout_epoch_ptr();
  B2C_REQUIRE(e != nullptr, "b2c_dropout_epoch_set: could not allocate the epoch word");
  B2C_CUDA(launch_k(epoch_set_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), e, value));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_dropout_epoch_use(const uint64_t* epoch_word) {
  g_epoch_use = epoch_word;
  return 0;
}

extern "C" float b2c_dropout_keep_scale(float p) {      // mirrors make_drop(): 8-bit keep decision per element
  if (!(p > 0.f)) return 1.f;
  uint32_t t16 = static_cast<uint32_t>(p * 65536.0 + 0.5);
  if (t16 > 65535u) t16 = 65535u;
  return 256.f / (256.f - static_cast<float>(t16 >> 8));
}

extern "C" int b2c_dropout_export_mask(int kind, float drop_p, uint64_t drop_seed, int64_t rows, int64_t cols, float* out,
                                       void* stream) {
  B2C_REQUIRE(out && rows > 0 && cols > 0, "b2c_dropout_export_mask: bad argument");
  B2C_REQUIRE(kind == 0 || kind == 1, "b2c_dropout_export_mask: kind must be 0 (elementwise) or 1 (attention)");
  B2C_REQUIRE(kind == 1 || cols % 4 == 0, "b2c_dropout_export_mask: elementwise sites have cols % 4 == 0");
  B2C_REQUIRE(drop_p > 0.f, "b2c_dropout_export_mask: drop_p must be > 0");
  const int64_t n4 = rows * ((cols + 3) / 4);
  const int grid = static_cast<int>(n4 / 256 + 1 < 4 * 148 ? n4 / 256 + 1 : 4 * 148);
  B2C_CUDA(launch_k(export_mask_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), kind, make_drop(drop_p, drop_seed), rows, cols, out));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
