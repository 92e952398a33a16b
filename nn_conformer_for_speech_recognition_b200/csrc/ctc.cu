// Fused log-softmax + CTC loss + gradient w.r.t. the LOGITS (sm_100a; HBM / latency bound, no tensor-core work).
//
// Replaces, in the reference's training step: `nn.LogSoftmax(-1)` at the end of ASRNN.forward
// (/root/reference/lib/standard/asrnn.py:44,254-256) followed by `CTCLoss(blank=hp.blank_idx, zero_infinity=True)`
// (/root/reference/lib/standard/runner.py:35,143, reduction 'mean': each utterance's negative log likelihood is divided by
// its target length (clamped to >= 1), then averaged over the batch; infinite losses and their gradients are zeroed).
// ATen runs that pair as log_softmax fwd/bwd over [B*T, V] plus three CTC kernels that walk the T steps with the whole
// vocabulary in tow (~1.1 ms of the 16.8 ms Conformer-M step, round-1 ncu launch list).  Here:
//
//   ctc_rows_kernel   one warp per (b, t < len_b) row: lse = logsumexp(logits[row, :]) and the log-probabilities of the
//                     S' = 2*S_b + 1 extended labels (blank, l1, blank, l2, ...) gathered into a compact [B, T, SP] buffer
//                     -- the only pass that reads the full logits in the forward direction (V = 1024: 98 MB for 64 x 373)
//   ctc_lattice_kernel  one CTA per utterance: the alpha recursion (threads [0, SP)) and the beta recursion (threads
//                     [SP, 2 SP)) run CONCURRENTLY, each on its own ping-pong row in shared memory, one named barrier per
//                     time step and half; log_alpha / log_beta rows go to global (L2-resident, B*T*SP*4 bytes each).  Then
//                     nll_b = -logaddexp(alpha[T-1, S'-1], alpha[T-1, S'-2]) and the state posteriors
//                     occ[t, s] = exp(alpha + beta - lp + nll) overwrite the alpha buffer.
//   ctc_grad_kernel   one warp per row: grad[row, c] = (softmax[row, c] - sum_{s: l'_s = c} occ[t, s]) * g / (B * max(S_b, 1)),
//                     zero for t >= len_b and for utterances whose loss is infinite (zero_infinity); g is READ FROM DEVICE
//                     MEMORY (the incoming gradient of the scalar loss), so nothing synchronises with the host.
// Conventions follow torch.nn.functional.ctc_loss: alpha and beta both include the emission at t, so their product counts
// lp[t, s] twice (hence "- lp"); a skip s -> s+2 is allowed when l'_{s+2} is not blank and differs from l'_s.
#include <math.h>

#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

constexpr float CTC_NEG_INF = -INFINITY;

__device__ __forceinline__ float log_add(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == CTC_NEG_INF) return CTC_NEG_INF;
  return m + log1pf(__expf(-fabsf(a - b)));
}
__device__ __forceinline__ float log_add3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == CTC_NEG_INF) return CTC_NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

struct CtcParams {
  const float* logits;       // [B, T, V]
  const int32_t* targets;    // [B, Smax]
  const int32_t* in_lens;    // [B]
  const int32_t* tgt_lens;   // [B]
  float* lse;                // [B, T]
  float* lpc;                // [B, T, SP]   log-probabilities of the extended labels
  float* la;                 // [B, T, SP]   log_alpha, later the state posteriors
  float* lb;                 // [B, T, SP]   log_beta
  float* nll;                // [B]
  float* loss;               // [1] (zeroed by the launcher)
  float* grad;               // [B, T, V]
  const float* grad_out;     // [1] device scalar, or nullptr = 1
  int B, T, V, Smax, SP, blank, zero_infinity;
  int stage_lp;              // 1: the lattice kernel stages an utterance's emissions in shared memory
};

__device__ __forceinline__ int ext_label(const int32_t* tgt, int s, int blank) { return (s & 1) ? tgt[s >> 1] : blank; }

// ---- pass 1: row log-sum-exp + gather of the extended labels' log-probabilities --------------------------------------
__global__ void __launch_bounds__(256) ctc_rows_kernel(const CtcParams p) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= static_cast<int64_t>(p.B) * p.T) return;
  const int b = static_cast<int>(row / p.T), t = static_cast<int>(row % p.T);
  const int len = min(p.in_lens[b], p.T);
  if (t >= len) return;
  const float* x = p.logits + row * p.V;
  float m = CTC_NEG_INF;
  const int V4 = p.V >> 2;
  for (int i = lane; i < V4; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(x + 4 * i);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int i = (V4 << 2) + lane; i < p.V; i += 32) m = fmaxf(m, x[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < V4; i += 32) {                       // second sweep hits L1/L2: the row was just read
    const float4 v = *reinterpret_cast<const float4*>(x + 4 * i);
    s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  for (int i = (V4 << 2) + lane; i < p.V; i += 32) s += __expf(x[i] - m);
  s = warp_sum(s);
  const float lse = m + __logf(s);
  if (lane == 0) p.lse[row] = lse;
  const int S1 = 2 * min(p.tgt_lens[b], p.Smax) + 1;
  const int32_t* tgt = p.targets + static_cast<int64_t>(b) * p.Smax;
  for (int sidx = lane; sidx < S1; sidx += 32)
    p.lpc[row * p.SP + sidx] = x[min(max(ext_label(tgt, sidx, p.blank), 0), p.V - 1)] - lse;
}

// ---- pass 2: alpha and beta recursions of one utterance, concurrently ------------------------------------------------
__global__ void __launch_bounds__(1024) ctc_lattice_kernel(const CtcParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];                       // [2 halves][2 rows][SP]
  const int b = blockIdx.x;
  const int SP = p.SP;
  const bool rec = static_cast<int>(threadIdx.x) < 2 * SP;   // recursion threads; the rest only help with the bulk phases
  const int half = (rec && static_cast<int>(threadIdx.x) >= SP) ? 1 : 0;         // 0: alpha threads, 1: beta threads
  const int s = rec ? static_cast<int>(threadIdx.x) - half * SP : SP;          // s = SP: never live
  const int len = min(p.in_lens[b], p.T);
  const int S = min(p.tgt_lens[b], p.Smax);
  const int S1 = 2 * S + 1;
  const int32_t* tgt = p.targets + static_cast<int64_t>(b) * p.Smax;
  const bool live = s < S1;
  const int lab = live ? ext_label(tgt, s, p.blank) : p.blank;
  // alpha: may state s be entered from s-2?   beta: may state s jump to s+2?
  bool skip = false;
  if (live && (s & 1)) {
    if (!half) skip = s >= 2 && tgt[(s >> 1) - 1] != lab;
    else skip = s + 2 < S1 && tgt[(s >> 1) + 1] != lab;
  }
  float* buf = sm + half * 2 * SP;
  // The emissions of this utterance ([len][SP] floats, 95 KB at T = 373, S' <= 64) are staged in shared memory by ONE
  // cooperative, coalesced sweep: read from global inside the recursion, every time step exposed a full memory latency
  // (~1 k cycles cold) to a loop whose arithmetic is ~100 cycles.  `stage` is false when they do not fit.
  float* slp = sm + 4 * SP;
  const bool stage = p.stage_lp != 0;
  const float* lpg = p.lpc + static_cast<int64_t>(b) * p.T * SP;
  if (stage) {
    const float4* src = reinterpret_cast<const float4*>(lpg);
    float4* dst4 = reinterpret_cast<float4*>(slp);
#pragma unroll 4
    for (int i = threadIdx.x; i < len * SP / 4; i += blockDim.x) dst4[i] = src[i];
    __syncthreads();
  }
  const float* lp = (stage ? slp : lpg) + s;
  float* dst = (half ? p.lb : p.la) + static_cast<int64_t>(b) * p.T * SP + s;
  const int bar_id = 1 + half;
  const int nthr = (SP + 31) & ~31;                   // threads of one half, rounded up to whole warps (SP % 32 == 0)
  if (len > 0 && rec) {
    // t0 = first step of this direction
    const int t0 = half ? len - 1 : 0;
    float v = CTC_NEG_INF;
    if (live) {
      const float e = lp[static_cast<int64_t>(t0) * SP];
      if (!half) { if (s <= 1) v = e; }                                  // alpha_0: blank or the first label
      else { if (s == S1 - 1 || s == S1 - 2) v = e; }                    // beta_{T-1}: the last blank or the last label
      dst[static_cast<int64_t>(t0) * SP] = v;
    }
    buf[s] = v;
    float e_next = 0.f;
    if (live && len > 1) e_next = lp[static_cast<int64_t>(half ? len - 2 : 1) * SP];
    for (int i = 1; i < len; ++i) {
      named_barrier_sync(bar_id, nthr);
      const int t = half ? len - 1 - i : i;
      const float* prev = buf + ((i - 1) & 1) * SP;
      float* cur = buf + (i & 1) * SP;
      const float e = e_next;
      if (live && i + 1 < len) e_next = lp[static_cast<int64_t>(half ? t - 1 : t + 1) * SP];   // the next step's emission
      float nv = CTC_NEG_INF;
      if (live) {
        const float a0 = prev[s];
        float a1 = CTC_NEG_INF, a2 = CTC_NEG_INF;
        if (!half) {
          if (s >= 1) a1 = prev[s - 1];
          if (skip) a2 = prev[s - 2];
        } else {
          if (s + 1 < S1) a1 = prev[s + 1];
          if (skip) a2 = prev[s + 2];
        }
        nv = log_add3(a0, a1, a2) + e;
        dst[static_cast<int64_t>(t) * SP] = nv;
      }
      cur[s] = nv;
    }
  }
  __syncthreads();
  // nll from the last alpha row (it sits in the alpha half's ping-pong buffer)
  float nll;
  if (len > 0) {
    const float* last = sm + ((len - 1) & 1) * SP;
    const float l1 = last[S1 - 1], l2 = S1 >= 2 ? last[S1 - 2] : CTC_NEG_INF;
    nll = -log_add(l1, l2);
  } else {
    nll = S == 0 ? 0.f : INFINITY;                   // no frames: only the empty target is reachable
  }
  const bool inf = !(nll < INFINITY);                // +inf or NaN
  if (threadIdx.x == 0) {
    p.nll[b] = nll;
    const float contrib = (inf && p.zero_infinity) ? 0.f : nll / static_cast<float>(max(S, 1)) / static_cast<float>(p.B);
    atomicAdd(p.loss, contrib);
  }
  // state posteriors occ[t, s] = exp(alpha + beta - lp + nll), written over log_alpha
  float* la = p.la + static_cast<int64_t>(b) * p.T * SP;
  const float* lb = p.lb + static_cast<int64_t>(b) * p.T * SP;
  const float* lpc = p.lpc + static_cast<int64_t>(b) * p.T * SP;
  // bulk phase, all threads, four independent load triples in flight per thread (the buffers are L2-resident)
  const int ntot = len * SP;
  for (int i0 = threadIdx.x; i0 < ntot; i0 += 4 * blockDim.x) {
    float a[4], bb[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = min(i0 + q * static_cast<int>(blockDim.x), ntot - 1);
      a[q] = la[idx]; bb[q] = lb[idx]; l[q] = stage ? slp[idx] : lpc[idx];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = i0 + q * static_cast<int>(blockDim.x);
      if (idx < ntot) {
        float o = 0.f;
        if ((idx % SP) < S1 && !inf && a[q] > CTC_NEG_INF && bb[q] > CTC_NEG_INF) o = __expf(a[q] + bb[q] - l[q] + nll);
        la[idx] = o;
      }
    }
  }
}

// ---- pass 3: gradient w.r.t. the logits ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctc_grad_kernel(const CtcParams p) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= static_cast<int64_t>(p.B) * p.T) return;
  const int b = static_cast<int>(row / p.T), t = static_cast<int>(row % p.T);
  const int len = min(p.in_lens[b], p.T);
  const int S = min(p.tgt_lens[b], p.Smax);
  float* g = p.grad + row * p.V;
  const int V4 = p.V >> 2;
  const float nll = p.nll[b];
  const bool dead = t >= len || (p.zero_infinity && !(nll < INFINITY));
  if (dead) {
    for (int i = lane; i < V4; i += 32) *reinterpret_cast<float4*>(g + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = (V4 << 2) + lane; i < p.V; i += 32) g[i] = 0.f;
    return;
  }
  const float scale = (p.grad_out ? *p.grad_out : 1.f) / (static_cast<float>(p.B) * static_cast<float>(max(S, 1)));
  const float* x = p.logits + row * p.V;
  const float lse = p.lse[row];
  for (int i = lane; i < V4; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(x + 4 * i);
    *reinterpret_cast<float4*>(g + 4 * i) = make_float4(__expf(v.x - lse) * scale, __expf(v.y - lse) * scale,
                                                        __expf(v.z - lse) * scale, __expf(v.w - lse) * scale);
  }
  for (int i = (V4 << 2) + lane; i < p.V; i += 32) g[i] = __expf(x[i] - lse) * scale;
  __syncwarp();
  __threadfence_block();
  // subtract the state posteriors: states that share a class (every blank; repeated labels) add up
  const int S1 = 2 * S + 1;
  const int32_t* tgt = p.targets + static_cast<int64_t>(b) * p.Smax;
  const float* occ = p.la + row * p.SP;
  for (int sidx = lane; sidx < S1; sidx += 32) {
    const float o = occ[sidx];
    if (o != 0.f) atomicAdd(g + min(max(ext_label(tgt, sidx, p.blank), 0), p.V - 1), -o * scale);
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" size_t b2c_ctc_workspace(int B, int T, int Smax) {
  const size_t SP = static_cast<size_t>((2 * Smax + 1 + 31) & ~31);
  const size_t rows = static_cast<size_t>(B) * T, rows4 = (rows + 3) & ~static_cast<size_t>(3);   // keeps lpc / la / lb 16-byte aligned
  return (rows4 + rows * 3 * SP + B) * sizeof(float);
}

// phases: 1 = forward (row pass + lattice: loss, nll, state posteriors kept in the workspace), 2 = gradient pass only (the
// workspace of an earlier phase-1 call on the same arguments), 3 = both.  grad_out: DEVICE scalar (incoming gradient of the
// loss) or NULL = 1.
extern "C" int b2c_ctc_loss(const float* logits, const int32_t* targets, const int32_t* input_lens,
                            const int32_t* target_lens, int B, int T, int V, int Smax, int blank, int zero_infinity,
                            int phases, const float* grad_out, float* loss, float* nll, float* grad_logits, void* workspace,
                            size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(logits && targets && input_lens && target_lens && loss && workspace, "b2c_ctc_loss: null pointer");
  B2C_REQUIRE(B > 0 && T > 0 && V > 0 && Smax >= 1, "b2c_ctc_loss: empty problem");
  B2C_REQUIRE(phases >= 1 && phases <= 3, "b2c_ctc_loss: phases must be 1, 2 or 3");
  B2C_REQUIRE(!(phases & 2) || grad_logits, "b2c_ctc_loss: the gradient pass needs grad_logits");
  B2C_REQUIRE(blank >= 0 && blank < V, "b2c_ctc_loss: blank index %d outside the vocabulary of %d", blank, V);
  B2C_REQUIRE(V % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                  (!grad_logits || (reinterpret_cast<uintptr_t>(grad_logits) & 15) == 0),
              "b2c_ctc_loss: V must be a multiple of 4 and logits / grad 16-byte aligned");
  const int SP = (2 * Smax + 1 + 31) & ~31;
  B2C_REQUIRE(2 * SP <= 1024, "b2c_ctc_loss: targets longer than 255 labels are not supported");
  B2C_REQUIRE(workspace_bytes >= b2c_ctc_workspace(B, T, Smax), "b2c_ctc_loss: workspace too small");
  B2C_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "b2c_ctc_loss: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CtcParams p;
  float* w = static_cast<float*>(workspace);
  const size_t rows = static_cast<size_t>(B) * T;
  p.logits = logits; p.targets = targets; p.in_lens = input_lens; p.tgt_lens = target_lens;
  p.lse = w; p.lpc = w + ((rows + 3) & ~static_cast<size_t>(3)); p.la = p.lpc + rows * SP; p.lb = p.la + rows * SP;
  p.nll = nll ? nll : p.lb + rows * SP;
  p.loss = loss; p.grad = grad_logits; p.grad_out = grad_out;
  p.B = B; p.T = T; p.V = V; p.Smax = Smax; p.SP = SP; p.blank = blank; p.zero_infinity = zero_infinity;
  p.stage_lp = 0;
  const unsigned row_blocks = static_cast<unsigned>((rows + 7) / 8);
  if (phases & 1) {
    B2C_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    B2C_CUDA(launch_k(ctc_rows_kernel, dim3(row_blocks), dim3(256), 0, st, p));
    size_t lat_smem = 4 * SP * sizeof(float);
    const size_t want = lat_smem + static_cast<size_t>(T) * SP * sizeof(float);
    p.stage_lp = want <= 200 * 1024 ? 1 : 0;
    if (p.stage_lp) lat_smem = want;
    static size_t configured = 0;
    if (lat_smem > 48 * 1024 && lat_smem > configured) {
      B2C_CUDA(cudaFuncSetAttribute(ctc_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 4096));
      configured = 200 * 1024 + 4096;
    }
    B2C_CUDA(launch_k(ctc_lattice_kernel, dim3(B), dim3(2 * SP > 512 ? 2 * SP : 512), lat_smem, st, p));
  }
  if (phases & 2) B2C_CUDA(launch_k(ctc_grad_kernel, dim3(row_blocks), dim3(256), 0, st, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
