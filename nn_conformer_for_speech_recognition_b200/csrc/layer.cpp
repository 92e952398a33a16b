// One C call per Conformer layer (eval mode): enqueues the layer's 20 kernel launches on the caller's stream.
//
// torchaudio ConformerLayer.forward, convolution_first=False (conformer.py:176-212), with dropout inactive and
// BatchNorm on its running statistics:
//     x += 0.5 * W2 SiLU(W1 LN(x))          x += Wo Attn(Wqkv LN(x))
//     x += Wpw2 SiLU(BN(DW(GLU(Wpw1 LN(x)))))    x += 0.5 * FFN2(x)     x = LN(x)
// Exists for the eager inference / NST pseudo-labelling path: issued from Python one launch at a time, a layer costs
// ~15 ctypes calls plus tensor bookkeeping, i.e. more host time than the GPU needs for the layer.  Composition only:
// every launch is one of the entry points above, on the same buffers the Python pipeline uses.
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#include "b200conformer.h"
#include "common.cuh"

static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

extern "C" size_t b2c_conformer_layer_eval_workspace(int B, int T, int d, int f, int H) {
  const size_t M = static_cast<size_t>(B) * T, Hd = static_cast<size_t>(H) * 64;
  return align256(M * d * 2) /*a*/ + align256(M * f * 2) /*h*/ + align256(M * 3 * Hd * 2) /*qkv*/ +
         align256(M * Hd * 2) /*att*/ + align256(M * 2 * d * 2) /*u*/ + align256(M * d * 2) /*c*/;
}

extern "C" int b2c_conformer_layer_eval(float* x, const b2c_layer_weights* w, const int32_t* lengths, int B, int T, int d,
                                        int f, int H, int head_dim, int k, float bn_eps, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  B2C_REQUIRE(x && w && lengths && workspace, "b2c_conformer_layer_eval: null pointer");
  B2C_REQUIRE(B > 0 && T > 0 && d > 0 && f > 0 && H > 0 && head_dim > 0 && head_dim <= 64,
              "b2c_conformer_layer_eval: bad dimensions");
  B2C_REQUIRE(workspace_bytes >= b2c_conformer_layer_eval_workspace(B, T, d, f, H),
              "b2c_conformer_layer_eval: workspace too small (%zu bytes)", workspace_bytes);
  const int64_t M = static_cast<int64_t>(B) * T, Hd = static_cast<int64_t>(H) * 64;
  char* p = static_cast<char*>(workspace);
  void* a = p;    p += align256(M * d * 2);
  void* h = p;    p += align256(M * f * 2);
  void* qkv = p;  p += align256(M * 3 * Hd * 2);
  void* att = p;  p += align256(M * Hd * 2);
  void* u = p;    p += align256(M * 2 * d * 2);
  void* c = p;
  const float eps = 1e-5f;
  int rc;
#define B2C_TRY(call) do { rc = (call); if (rc) return rc; } while (0)
  auto gemm = [&](const void* A, int64_t lda, const void* Wt, int64_t N, int64_t K, const float* bias, int act,
                  const float* resid, void* out, int out_f32, float alpha) -> int {
    b2c_epilogue ep = {};
    ep.bias = bias; ep.resid = resid; ep.out = out; ep.ld_out = N; ep.alpha = alpha; ep.act = act;
    ep.out_dtype = out_f32 ? B2C_F32 : B2C_BF16;
    return b2c_gemm_bf16(A, lda, 0, Wt, K, 0, M, N, K, 1, 0, &ep, stream);
  };
  // opt-in (B2C_FFN_FUSED=1): measured slower than LN + 2 GEMMs at the benchmark shapes, see csrc/ffn_fused.cu / DESIGN.md
  static const bool fused_on = [] { const char* e = getenv("B2C_FFN_FUSED"); return e && e[0] == '1'; }();
  const bool fuse_ffn = fused_on && d % 16 == 0 && d <= 256 && f % 8 == 0;
  auto ffn = [&](const b2c_ffn_weights& q) -> int {
    int r;
    if (fuse_ffn)      // LN -> GEMM -> SiLU -> GEMM -> 1/2-residual in one kernel: the hidden activation stays on chip
      return b2c_ffn_fused_eval(x, x, q.ln_w, q.ln_b, q.w1, q.b1, q.w2, q.b2, M, d, f, 0.5f, eps, stream);
    if ((r = b2c_layernorm_fwd(x, q.ln_w, q.ln_b, a, B2C_BF16, nullptr, nullptr, M, d, eps, stream))) return r;
    if ((r = gemm(a, d, q.w1, f, d, q.b1, B2C_ACT_SILU, nullptr, h, 0, 1.f))) return r;
    return gemm(h, f, q.w2, d, f, q.b2, B2C_ACT_NONE, x, x, 1, 0.5f);
  };
  // 1. half-step FFN
  B2C_TRY(ffn(w->ffn1));
  // 2. self-attention (heads zero-padded to 64 in in_w / in_b / out_w by the caller)
  B2C_TRY(b2c_layernorm_fwd(x, w->att_ln_w, w->att_ln_b, a, B2C_BF16, nullptr, nullptr, M, d, eps, stream));
  B2C_TRY(gemm(a, d, w->in_w, 3 * Hd, d, w->in_b, B2C_ACT_NONE, nullptr, qkv, 0, 1.f));
  B2C_TRY(b2c_attention_fwd(qkv, 3 * Hd, Hd, 2 * Hd, lengths, att, Hd, nullptr, B, T, H,
                            1.f / sqrtf(static_cast<float>(head_dim)), 0.f, 0, stream));
  B2C_TRY(gemm(att, Hd, w->out_w, d, Hd, w->out_b, B2C_ACT_NONE, x, x, 1, 1.f));
  // 3. convolution module
  B2C_TRY(b2c_layernorm_fwd(x, w->conv_ln_w, w->conv_ln_b, a, B2C_BF16, nullptr, nullptr, M, d, eps, stream));
  B2C_TRY(gemm(a, d, w->pw1_w, 2 * d, d, w->pw1_b, B2C_ACT_NONE, nullptr, u, 0, 1.f));
  B2C_TRY(b2c_convmod_eval_fwd(u, w->dw_w, w->dw_b, w->bn_w, w->bn_b, w->bn_mean, w->bn_var, bn_eps, c, B, T, d, k, stream));
  B2C_TRY(gemm(c, d, w->pw2_w, d, d, w->pw2_b, B2C_ACT_NONE, x, x, 1, 1.f));
  // 4. half-step FFN, 5. final LayerNorm (in place: a warp holds its whole row in registers before it writes)
  B2C_TRY(ffn(w->ffn2));
  B2C_TRY(b2c_layernorm_fwd(x, w->final_ln_w, w->final_ln_b, x, B2C_F32, nullptr, nullptr, M, d, eps, stream));
#undef B2C_TRY
  return 0;
}
