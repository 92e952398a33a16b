// Input pipeline of the reference on the GPU: log-mel extraction and batch padding.
//
//  b2c_logmel         librosa.feature.melspectrogram(y, sr, n_mels) -> log -> min-max normalisation, exactly the steps of
//                     `SpeechCommands.get_mels` (/root/reference/lib/standard/speechcommands.py:113-119):
//                        S   = |STFT(y, n_fft, hop, hann, center=True, zero padding)|^2          (librosa defaults 2048 / 512)
//                        mel = basis[n_mels, n_fft/2+1] @ S                                       (Slaney scale and norm; host-built)
//                        mel = where(mel < 1e-10, 0, log(mel));  mel -= min(mel);  mel /= max(mel)
//                     written straight into the zero-padded batch tensor the model consumes ([B, n_mels, T_max]), so the
//                     reference's python-list padding (speechcommands.py:186-190) disappears for the waveform path.
//                     The reference publishes 122-125 clips/s for this step on the Colab host (main.ipynb:407-410).
//  b2c_pad_batch      the padding of `SpeechCommands.get_batch` (speechcommands.py:174-197) for pre-computed features:
//                     ragged utterances packed back to back -> [B, F, T_max] zero-padded, rows >= n all zero.
// FFT: one CTA per frame, radix-2 decimation-in-time on 2^k complex points in shared memory (the frame is real: its
// imaginary parts are zero; n_fft = 2048 costs 11 passes of 1024 butterflies -- the front end is a vanishing fraction of
// the encoder's work, so the straightforward transform is used), twiddles by sincospif.
#include <math.h>

#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

struct LogmelParams {
  const float* wave;          // packed waveforms
  const int64_t* wave_off;    // [n + 1] sample offsets
  const float* basis;         // [n_mels, n_fft/2 + 1]
  const int32_t* f_lo;        // [n_mels] first / one-past-last non-zero bin of each filter
  const int32_t* f_hi;
  float* mel;                 // [n, n_mels, frames_max] scratch (power mel spectrogram)
  float* out;                 // [B, n_mels, t_max]
  int32_t* frames_out;        // [B] frames per utterance (0 for padding rows)
  int n, B, n_fft, log2n, hop, n_mels, frames_max, t_max;
};

__global__ void __launch_bounds__(256) logmel_power_kernel(const LogmelParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float2 sx[];                         // [n_fft]
  const int u = blockIdx.y, fr = blockIdx.x;
  const int64_t o0 = p.wave_off[u], len = p.wave_off[u + 1] - o0;
  const int frames = static_cast<int>(len / p.hop) + 1;  // librosa center=True: 1 + len // hop
  if (fr >= frames) return;
  const int N = p.n_fft, half = N >> 1;
  const float* y = p.wave + o0;
  const int64_t start = static_cast<int64_t>(fr) * p.hop - half;
  // windowed frame, written in bit-reversed order
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t t = start + i;
    const float s = (t >= 0 && t < len) ? y[t] : 0.f;    // pad_mode = "constant"
    const float w = 0.5f - 0.5f * cospif(2.0f * static_cast<float>(i) / static_cast<float>(N));   // periodic Hann
    const int r = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - p.log2n));
    sx[r] = make_float2(s * w, 0.f);
  }
  __syncthreads();
  for (int st = 1; st <= p.log2n; ++st) {
    const int m = 1 << st, mh = m >> 1;
    for (int b = threadIdx.x; b < half; b += blockDim.x) {
      const int k = b & (mh - 1);
      const int i0 = ((b >> (st - 1)) << st) + k, i1 = i0 + mh;
      float sn, cs;
      sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(m), &sn, &cs);
      const float2 a = sx[i0], c = sx[i1];
      const float2 t = make_float2(c.x * cs - c.y * sn, c.x * sn + c.y * cs);
      sx[i0] = make_float2(a.x + t.x, a.y + t.y);
      sx[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
  // power spectrum into the real parts of the first half+1 entries
  for (int k = threadIdx.x; k <= half; k += blockDim.x) {
    const float2 v = sx[k];
    sx[k].x = v.x * v.x + v.y * v.y;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int nb = half + 1;
  for (int m = warp; m < p.n_mels; m += nw) {
    float acc = 0.f;
    for (int k = p.f_lo[m] + lane; k < p.f_hi[m]; k += 32) acc = fmaf(__ldg(p.basis + static_cast<int64_t>(m) * nb + k), sx[k].x, acc);
    acc = warp_sum(acc);
    if (lane == 0) p.mel[(static_cast<int64_t>(u) * p.n_mels + m) * p.frames_max + fr] = acc;
  }
}

// one CTA per batch row: log, min, max, normalise, zero padding
__global__ void __launch_bounds__(256) logmel_norm_kernel(const LogmelParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][8];
  const int u = blockIdx.x;
  float* out = p.out + static_cast<int64_t>(u) * p.n_mels * p.t_max;
  const int64_t total = static_cast<int64_t>(p.n_mels) * p.t_max;
  int frames = 0;
  if (u < p.n) frames = static_cast<int>((p.wave_off[u + 1] - p.wave_off[u]) / p.hop) + 1;
  if (frames > p.t_max) frames = p.t_max;
  if (threadIdx.x == 0) p.frames_out[u] = frames;
  const float* mel = p.mel + static_cast<int64_t>(u) * p.n_mels * p.frames_max;
  const int64_t nval = static_cast<int64_t>(p.n_mels) * frames;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < nval; i += blockDim.x) {
    const int m = static_cast<int>(i / frames), t = static_cast<int>(i % frames);
    const float v = mel[static_cast<int64_t>(m) * p.frames_max + t];
    const float l = v < 1e-10f ? 0.f : logf(v);                                  // speechcommands.py:114
    mn = fminf(mn, l);
    mx = fmaxf(mx, l);
  }
  mn = -warp_max(-mn);
  mx = warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = mn; red[1][warp] = mx; }
  __syncthreads();
  mn = red[0][0]; mx = red[1][0];
  for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[0][w]); mx = fmaxf(mx, red[1][w]); }
  const float inv = 1.f / (mx - mn);                                              // speechcommands.py:117-119
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int m = static_cast<int>(i / p.t_max), t = static_cast<int>(i % p.t_max);
    float o = 0.f;
    if (t < frames) {
      const float v = mel[static_cast<int64_t>(m) * p.frames_max + t];
      o = ((v < 1e-10f ? 0.f : logf(v)) - mn) * inv;
    }
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) pad_batch_kernel(const float* __restrict__ packed, const int64_t* __restrict__ off,
                                                       const int32_t* __restrict__ lens, float* __restrict__ out, int n, int F,
                                                       int t_max) {
  pdl_trigger();
  pdl_wait();
  const int u = blockIdx.y;
  const int len = u < n ? min(lens[u], t_max) : 0;
  const int src_len = u < n ? lens[u] : 0;
  const float* src = u < n ? packed + off[u] : packed;
  float* dst = out + static_cast<int64_t>(u) * F * t_max;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<int64_t>(F) * t_max;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i / t_max), t = static_cast<int>(i % t_max);
    dst[i] = t < len ? src[static_cast<int64_t>(f) * src_len + t] : 0.f;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_logmel(const float* wave, const int64_t* wave_off, int n, int B, const float* basis, const int32_t* f_lo,
                          const int32_t* f_hi, int n_fft, int hop, int n_mels, int frames_max, int t_max, float* mel_scratch,
                          float* out, int32_t* frames_out, void* stream) {
  B2C_REQUIRE(wave && wave_off && basis && f_lo && f_hi && mel_scratch && out && frames_out, "b2c_logmel: null pointer");
  B2C_REQUIRE(n >= 0 && B >= n && B > 0 && n_mels > 0 && frames_max > 0 && t_max > 0 && hop > 0, "b2c_logmel: bad sizes");
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  B2C_REQUIRE((1 << log2n) == n_fft && n_fft >= 64 && n_fft <= 4096, "b2c_logmel: n_fft must be a power of two in [64, 4096]");
  LogmelParams p;
  p.wave = wave; p.wave_off = wave_off; p.basis = basis; p.f_lo = f_lo; p.f_hi = f_hi; p.mel = mel_scratch; p.out = out;
  p.frames_out = frames_out;
  p.n = n; p.B = B; p.n_fft = n_fft; p.log2n = log2n; p.hop = hop; p.n_mels = n_mels; p.frames_max = frames_max; p.t_max = t_max;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n > 0)
    B2C_CUDA(launch_k(logmel_power_kernel, dim3(frames_max, n), dim3(256), n_fft * sizeof(float2), st, p));
  B2C_CUDA(launch_k(logmel_norm_kernel, dim3(B), dim3(256), 0, st, p));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_pad_batch(const float* packed, const int64_t* off, const int32_t* lens, int n, int B, int F, int t_max,
                             float* out, void* stream) {
  B2C_REQUIRE(packed && off && lens && out, "b2c_pad_batch: null pointer");
  B2C_REQUIRE(n >= 0 && B >= n && B > 0 && F > 0 && t_max > 0, "b2c_pad_batch: bad sizes");
  const int64_t per = static_cast<int64_t>(F) * t_max;
  const int gx = static_cast<int>(per / 256 + 1 < 64 ? per / 256 + 1 : 64);
  B2C_CUDA(launch_k(pad_batch_kernel, dim3(gx, B), dim3(256), 0, static_cast<cudaStream_t>(stream), packed, off, lens, out, n, F, t_max));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
