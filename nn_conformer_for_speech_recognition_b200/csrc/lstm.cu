// LSTM cell kernels: the elementwise half of the reference's bi-LSTM "decoder"
//   self.lstm = nn.LSTM(projection_out_size, standard_decoder_nodes, bidirectional=True)   /root/reference/lib/standard/asrnn.py:38
//   x = self.lstm(x)[0]                                                                     asrnn.py:252
// (torch gate order i, f, g, o: torch/nn/modules/rnn.py LSTM).  The matrix half -- X W_ih^T for all steps at once, the
// per-step h_{t-1} W_hh^T, and in the backward pass dG W_hh, dG^T H, dG^T X, dG W_ih -- runs on the tcgen05 GEMM
// (b2c_gemm_bf16); these two kernels do, per time step, everything between two recurrent GEMMs in one pass:
//   lstm_cell_fwd   gates(pre) -> i,f,o = sigmoid, g = tanh; c = f c_prev + i g; h = o tanh(c); writes c (fp32), h as fp32 into
//                   the [L, N, dirs*H] output and as bf16 (operand of the next step's GEMM, and of the weight gradient), and
//                   the four activated gates (fp32) for the backward pass
//   lstm_cell_bwd   dh = dOut[t] + dh_rec;  dc = dc_next + dh o (1 - tanh(c)^2);  d(pre-activations) as bf16 [N, 4H] (operand
//                   of the recurrent dgrad GEMM and of all weight-gradient GEMMs), dc_prev = dc f
// Thread = 4 adjacent hidden units of one row (16-byte fp32 accesses, 8-byte bf16), memory bound: 12 fp32 + bf16 words per unit.
#include "b200conformer.h"
#include "common.cuh"

namespace b2c {

__device__ __forceinline__ float tanh_acc(float x) { return tanhf(x); }            // full-precision: errors here feed back for L steps
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                           float* __restrict__ c_out, float* __restrict__ h_out, int64_t ld_h,
                                                           __nv_bfloat16* __restrict__ h_bf16, float* __restrict__ acts,
                                                           int64_t N, int H) {
  pdl_trigger();
  pdl_wait();
  const int hq = H >> 2;
  const int64_t n4 = N * hq;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n4;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = idx / hq;
    const int j = static_cast<int>(idx - n * hq) * 4;
    const float* g = gates + n * 4 * H + j;
    const float4 pi = *reinterpret_cast<const float4*>(g), pf = *reinterpret_cast<const float4*>(g + H);
    const float4 pg = *reinterpret_cast<const float4*>(g + 2 * H), po = *reinterpret_cast<const float4*>(g + 3 * H);
    float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c_prev) cp = *reinterpret_cast<const float4*>(c_prev + n * H + j);
    const float vi[4] = {sigmoid_acc(pi.x), sigmoid_acc(pi.y), sigmoid_acc(pi.z), sigmoid_acc(pi.w)};
    const float vf[4] = {sigmoid_acc(pf.x), sigmoid_acc(pf.y), sigmoid_acc(pf.z), sigmoid_acc(pf.w)};
    const float vg[4] = {tanh_acc(pg.x), tanh_acc(pg.y), tanh_acc(pg.z), tanh_acc(pg.w)};
    const float vo[4] = {sigmoid_acc(po.x), sigmoid_acc(po.y), sigmoid_acc(po.z), sigmoid_acc(po.w)};
    const float c0[4] = {cp.x, cp.y, cp.z, cp.w};
    float c[4], h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      c[e] = fmaf(vf[e], c0[e], vi[e] * vg[e]);
      h[e] = vo[e] * tanh_acc(c[e]);
    }
    *reinterpret_cast<float4*>(c_out + n * H + j) = make_float4(c[0], c[1], c[2], c[3]);
    *reinterpret_cast<float4*>(h_out + n * ld_h + j) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint2*>(h_bf16 + n * H + j) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
    if (acts) {
      float* a = acts + n * 4 * H + j;
      *reinterpret_cast<float4*>(a) = make_float4(vi[0], vi[1], vi[2], vi[3]);
      *reinterpret_cast<float4*>(a + H) = make_float4(vf[0], vf[1], vf[2], vf[3]);
      *reinterpret_cast<float4*>(a + 2 * H) = make_float4(vg[0], vg[1], vg[2], vg[3]);
      *reinterpret_cast<float4*>(a + 3 * H) = make_float4(vo[0], vo[1], vo[2], vo[3]);
    }
  }
}

__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ dout, int64_t ld_dout,
                                                           const float* __restrict__ dh_rec, float* __restrict__ dc,
                                                           const float* __restrict__ acts, const float* __restrict__ c_prev,
                                                           const float* __restrict__ c_cur, __nv_bfloat16* __restrict__ dgates,
                                                           int64_t N, int H) {
  pdl_trigger();
  pdl_wait();
  const int hq = H >> 2;
  const int64_t n4 = N * hq;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n4;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = idx / hq;
    const int j = static_cast<int>(idx - n * hq) * 4;
    float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dout) dh = *reinterpret_cast<const float4*>(dout + n * ld_dout + j);
    if (dh_rec) {
      const float4 r = *reinterpret_cast<const float4*>(dh_rec + n * H + j);
      dh.x += r.x; dh.y += r.y; dh.z += r.z; dh.w += r.w;
    }
    const float* a = acts + n * 4 * H + j;
    const float4 ai = *reinterpret_cast<const float4*>(a), af = *reinterpret_cast<const float4*>(a + H);
    const float4 ag = *reinterpret_cast<const float4*>(a + 2 * H), ao = *reinterpret_cast<const float4*>(a + 3 * H);
    float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c_prev) cp = *reinterpret_cast<const float4*>(c_prev + n * H + j);
    const float4 cc = *reinterpret_cast<const float4*>(c_cur + n * H + j);
    const float4 dcn = *reinterpret_cast<const float4*>(dc + n * H + j);
    const float vdh[4] = {dh.x, dh.y, dh.z, dh.w}, vi[4] = {ai.x, ai.y, ai.z, ai.w}, vf[4] = {af.x, af.y, af.z, af.w};
    const float vg[4] = {ag.x, ag.y, ag.z, ag.w}, vo[4] = {ao.x, ao.y, ao.z, ao.w}, vcp[4] = {cp.x, cp.y, cp.z, cp.w};
    const float vc[4] = {cc.x, cc.y, cc.z, cc.w}, vdc[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
    float di[4], df[4], dg[4], dob[4], dcp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float tc = tanh_acc(vc[e]);
      const float dct = fmaf(vdh[e] * vo[e], 1.f - tc * tc, vdc[e]);
      dob[e] = vdh[e] * tc * vo[e] * (1.f - vo[e]);
      di[e] = dct * vg[e] * vi[e] * (1.f - vi[e]);
      df[e] = dct * vcp[e] * vf[e] * (1.f - vf[e]);
      dg[e] = dct * vi[e] * (1.f - vg[e] * vg[e]);
      dcp[e] = dct * vf[e];
    }
    *reinterpret_cast<float4*>(dc + n * H + j) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    __nv_bfloat16* o = dgates + n * 4 * H + j;
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(di[0], di[1]), pack_bf16(di[2], di[3]));
    *reinterpret_cast<uint2*>(o + H) = make_uint2(pack_bf16(df[0], df[1]), pack_bf16(df[2], df[3]));
    *reinterpret_cast<uint2*>(o + 2 * H) = make_uint2(pack_bf16(dg[0], dg[1]), pack_bf16(dg[2], dg[3]));
    *reinterpret_cast<uint2*>(o + 3 * H) = make_uint2(pack_bf16(dob[0], dob[1]), pack_bf16(dob[2], dob[3]));
  }
}

static int cell_grid(int64_t N, int H) {
  const int64_t n4 = N * (H >> 2);
  const int64_t g = (n4 + 255) / 256;
  const int64_t cap = 8ll * num_sms();
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_lstm_cell_fwd(const float* gates, const float* c_prev, float* c_out, float* h_out, int64_t ld_h, void* h_bf16,
                                 float* acts, int64_t N, int H, void* stream) {
  B2C_REQUIRE(gates && c_out && h_out && h_bf16, "b2c_lstm_cell_fwd: null pointer");
  B2C_REQUIRE(N > 0 && H > 0 && H % 4 == 0 && ld_h % 4 == 0 && ld_h >= H, "b2c_lstm_cell_fwd: N > 0, H %% 4 == 0, ld_h %% 4 == 0 and >= H");
  B2C_REQUIRE(((reinterpret_cast<uintptr_t>(gates) | reinterpret_cast<uintptr_t>(c_out) | reinterpret_cast<uintptr_t>(h_out) |
                reinterpret_cast<uintptr_t>(c_prev) | reinterpret_cast<uintptr_t>(acts)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(h_bf16) & 7) == 0,
              "b2c_lstm_cell_fwd: pointers must be 16-byte aligned");
  B2C_CUDA(launch_k(lstm_cell_fwd_kernel, dim3(cell_grid(N, H)), dim3(256), 0, static_cast<cudaStream_t>(stream), gates, c_prev, c_out,
                    h_out, ld_h, static_cast<__nv_bfloat16*>(h_bf16), acts, N, H));
  B2C_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b2c_lstm_cell_bwd(const float* dout, int64_t ld_dout, const float* dh_rec, float* dc, const float* acts,
                                 const float* c_prev, const float* c_cur, void* dgates, int64_t N, int H, void* stream) {
  B2C_REQUIRE(dc && acts && c_cur && dgates, "b2c_lstm_cell_bwd: null pointer");
  B2C_REQUIRE(N > 0 && H > 0 && H % 4 == 0 && ld_dout % 4 == 0, "b2c_lstm_cell_bwd: N > 0, H %% 4 == 0, ld_dout %% 4 == 0");
  B2C_REQUIRE(((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dh_rec) | reinterpret_cast<uintptr_t>(dc) |
                reinterpret_cast<uintptr_t>(acts) | reinterpret_cast<uintptr_t>(c_prev) | reinterpret_cast<uintptr_t>(c_cur)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dgates) & 7) == 0,
              "b2c_lstm_cell_bwd: pointers must be 16-byte aligned");
  B2C_CUDA(launch_k(lstm_cell_bwd_kernel, dim3(cell_grid(N, H)), dim3(256), 0, static_cast<cudaStream_t>(stream), dout, ld_dout, dh_rec,
                    dc, acts, c_prev, c_cur, static_cast<__nv_bfloat16*>(dgates), N, H));
  B2C_CUDA(cudaGetLastError());
  return 0;
}
