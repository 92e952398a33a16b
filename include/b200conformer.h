/* b200conformer.h -- C ABI of libb200conformer.so: the sm_100a kernels behind the Conformer-encoder hot path.
 *
 * The reference (icadriani/nn_conformer_for_speech_recognition) is pure Python and has no FFI layer
 * (SURVEY.md R1); its hot path is the set of torch modules listed beside each entry point below.  This header
 * is the drop-in boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub): plain device
 * pointers and sizes, no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - return 0 on success, <0 for a rejected argument, >0 = cudaError_t; b2c_last_error() explains (thread-local);
 *   - activations are token-major: row = b*T + t, features contiguous; bf16 unless stated; statistics fp32.
 */
#ifndef B200CONFORMER_H_
#define B200CONFORMER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2C_VERSION 1

enum { B2C_BF16 = 0, B2C_F32 = 1 };
enum { B2C_ACT_NONE = 0, B2C_ACT_SILU = 1, B2C_ACT_SILU_BWD = 2 };

const char* b2c_last_error(void);
int b2c_version(void);
/* Dropout (torchaudio conformer.py:74,106,108,152,153; asrnn.py:210).  Masks are a pure function of
 * (seed, element index, epoch word); the epoch word is read by the KERNELS, so a replayed CUDA graph still draws fresh
 * masks each step.  There is one step counter per device:
 *   _bump      counter += 1                                  (one tiny launch)
 *   _snapshot  counter += 1; *dst = counter                  dst = a device word owned by ONE training forward call
 *   _set       counter = value                               (tests: makes mask sequences reproducible)
 *   _use       host-side, thread-local: kernels launched by this thread from now on read `epoch_word` instead of the
 *              device counter (NULL restores the counter).  A forward installs its snapshot for its own launches and
 *              its backward installs the same word again, so interleaved forward/backward pairs regenerate the masks
 *              each forward really used.
 *   _export_mask  debug: out fp32 [rows, cols] = the scale (0 or 1/keep) the kernels apply at a site with this
 *              (drop_p, drop_seed) under the currently installed epoch word.  kind 0: elementwise sites (GEMM residual
 *              epilogues, b2c_scale_cast, b2c_layernorm_bwd_cast; element index = r*cols + c).  kind 1: attention
 *              probabilities, rows = (b*H + h)*T + q, cols = T keys. */
int b2c_dropout_epoch_bump(void* stream);
int b2c_dropout_epoch_snapshot(uint64_t* dst, void* stream);
int b2c_dropout_epoch_set(uint64_t value, void* stream);
int b2c_dropout_epoch_use(const uint64_t* epoch_word);
int b2c_dropout_export_mask(int kind, float drop_p, uint64_t drop_seed, int64_t rows, int64_t cols, float* out,
                            void* stream);
/* fails (returns <0) unless the current CUDA device is compute capability 10.x -- there is no fallback path */
int b2c_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * GEMM  D[M,N] = epi(A * B^T), bf16 x bf16 -> fp32 (tcgen05.mma, accumulators in TMEM, operands by TMA).
 * Replaces torch.nn.Linear / Conv1d(kernel 1) + the elementwise ops around them:
 *   FFN   torchaudio/models/conformer.py:102-109 (Linear+SiLU, Linear) and the 0.5*x+residual at :186-187,208-209
 *   MHSA  in/out projections, torch/nn/functional.py:5798,6690 (called from conformer.py:194-200)
 *   conv module pointwise convs, conformer.py:44-51,66-73
 * a_mn / b_mn: 0 = the operand is [rows][K] with K contiguous; 1 = it is stored [K][rows] (rows contiguous).
 *   forward (0,0) with B = weight (out,in);  dX = dY*W is (0,1) with B = W;  dW = dY^T*X is (1,1).
 * epilogue, in this order:  v = acc (+ bias[n]);  out2 = bf16(v) if out2;  v = act(v);  v *= dropout(seed, m*N+n);
 *                           v *= alpha;  then  out = bf16(v) | out = v + resid (fp32) | atomicAdd(out, v) (fp32);
 *                           colsum[n] += sum_m v  (optional, bf16 outputs).
 * ------------------------------------------------------------------------------------------------------- */
/* Value stored in `out2` (the saved pre-activation of act = SILU) where the activation itself is dropped by the fused
 * dropout: SiLU'(B2C_DROPPED_PREACT) == 0 exactly, so the backward dgrad (act = SILU_BWD, aux = that tensor,
 * alpha = 1 / keep-probability, see b2c_dropout_keep_scale) needs no mask regeneration. */
#define B2C_DROPPED_PREACT (-1.0e4f)
/* exact scale the fused dropout applies to kept elements for probability p (the keep decision uses 8 bits per element) */
float b2c_dropout_keep_scale(float p);

typedef struct b2c_epilogue {
  const void* bias;   /* fp32 [N] or NULL */
  const void* aux;    /* bf16 [M, ld_aux]: pre-activation for B2C_ACT_SILU_BWD, else NULL */
  const void* resid;  /* fp32 [M, ld_out] added to an fp32 output (may alias out), or NULL */
  void* out;          /* bf16 or fp32 [M, ld_out] */
  void* out2;         /* optional bf16 [M, ld_out2] copy of acc+bias (saved pre-activation) */
  int64_t ld_out, ld_out2, ld_aux;
  float alpha;
  float drop_p;       /* 0 = no dropout */
  uint64_t drop_seed;
  int act;            /* B2C_ACT_* */
  int out_dtype;      /* B2C_BF16 | B2C_F32 */
  int accumulate;     /* 1: fp32 atomicAdd into out (split-K / gradient accumulation) */
  int reserved;
  float* colsum;      /* optional fp32 [N]: += column sums of the values written to `out` (bias gradient of the next op) */
} b2c_epilogue;

int b2c_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                  int64_t K, int split_k, int block_n, const b2c_epilogue* ep, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused macaron feed-forward half-step, INFERENCE:  out = x + alpha * (W2 SiLU(W1 LayerNorm(x) + b1) + b2)
 * = torchaudio `_FeedForwardModule` (conformer.py:102-109, dropouts inactive) + the `0.5 * x + residual` of
 * `ConformerLayer.forward` (conformer.py:185-187,207-209) in ONE kernel: the [M, f] hidden activation stays in shared
 * memory / TMEM between the two tcgen05 products (SURVEY.md section 2a "fused LN->GEMM->SiLU->GEMM->1/2-residual").
 *   x, out fp32 [M, d] (may alias); ln_w, ln_b, b2 fp32 [d]; b1 fp32 [f]; w1 bf16 [f, d]; w2 bf16 [d, f] (row-major,
 *   torch Linear layout).  d % 16 == 0, d <= 256; f % 8 == 0.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_ffn_fused_eval(const float* x, float* out, const float* ln_w, const float* ln_b, const void* w1, const float* b1,
                       const void* w2, const float* b2, int64_t M, int d, int f, float alpha, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (biased variance, eps) -- torch.nn.LayerNorm at conformer.py:42,103,151,165.
 * x fp32 [rows,d] -> y (bf16 or fp32); mean/rstd fp32 [rows] optional (saved for backward).
 * bwd: dx (+)= LN'(dy); dgamma/dbeta fp32 [d] are ACCUMULATED (atomicAdd) -- zero them first.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype, float* mean,
                      float* rstd, int64_t rows, int d, float eps, void* stream);
int b2c_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean,
                      const float* rstd, float* dx, int dx_accumulate, float* dgamma, float* dbeta, int64_t rows, int d,
                      void* stream);
/* same, and additionally  cast_out (bf16 [rows,d]) = alpha * dropout(dx_new)  with colsum (fp32 [d], optional) += column
 * sums of cast_out: the operand and bias gradient the next sub-layer's backward starts from (fuses b2c_scale_cast, i.e.
 * the gradient of `x + alpha*Dropout(f(x))` at conformer.py:185-209, into this pass over dx). */
int b2c_layernorm_bwd_cast(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean,
                           const float* rstd, float* dx, int dx_accumulate, float* dgamma, float* dbeta, int64_t rows,
                           int d, void* cast_out, float alpha, float drop_p, uint64_t drop_seed, float* colsum,
                           void* stream);

/* Two LayerNorms back to back -- the final LayerNorm of Conformer layer i (torchaudio conformer.py:211) followed by the
 * LayerNorm of layer i+1's first feed-forward module (conformer.py:103,185) -- in one pass over the rows:
 *   _fwd_pair: y1 = LN1(x) (fp32, stored), a2 = LN2(y1) (bf16, stored); mean/rstd of both saved when non-null.
 *              Bit-identical to b2c_layernorm_fwd(x -> y1, fp32) followed by b2c_layernorm_fwd(y1 -> a2, bf16).
 *   _bwd_pair: dy1 = prev + LN2'(da2; y1);  dx = LN1'(dy1; x);  cast_out = bf16(alpha * dropout(dx)), colsum += its column
 *              sums; dgamma / dbeta of BOTH LayerNorms are accumulated (atomicAdd).  Equals b2c_layernorm_bwd(da2, y1, ...,
 *              dx = prev, accumulate = 1) followed by b2c_layernorm_bwd_cast(prev, fp32, x, ..., dx, accumulate = 0, ...)
 *              without writing / re-reading the intermediate gradient.  `prev` is only read. */
int b2c_layernorm_fwd_pair(const float* x, const float* gamma1, const float* beta1, float* y1, float* mean1, float* rstd1,
                           const float* gamma2, const float* beta2, void* a2, float* mean2, float* rstd2, int64_t rows,
                           int d, float eps, void* stream);
int b2c_layernorm_bwd_pair(const void* da2, const float* y1, const float* gamma2, const float* mean2, const float* rstd2,
                           float* dgamma2, float* dbeta2, const float* prev, const float* x, const float* gamma1,
                           const float* mean1, const float* rstd1, float* dgamma1, float* dbeta1, float* dx, int64_t rows,
                           int d, void* cast_out, float alpha, float drop_p, uint64_t drop_seed, float* colsum,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Conv-module chain  GLU -> depthwise conv1d('same', cross-correlation) -> BatchNorm1d -> SiLU
 * (conformer.py:52-65; frames past `lengths` are NOT zeroed, conformer.py:85-88).
 *   u   bf16 [B*T, 2d]  pointwise-conv-1 output (first d channels are gated by sigmoid of the last d)
 *   wdw fp32 [d,k] (= sequential.2.weight[:,0,:]), bdw fp32 [d]
 * eval : out = SiLU(BN_running(dw(GLU(u))))           one kernel
 * train: _dw writes y = dw(GLU(u)) (bf16) and accumulates per-channel sum / sum-of-squares (fp64 [2*d], zeroed
 *        by the caller); _bn_finalize turns them into mean/rstd, updates running stats (momentum, unbiased var);
 *        _bn_apply writes SiLU(BN_batch(y)).
 * ------------------------------------------------------------------------------------------------------- */
int b2c_convmod_eval_fwd(const void* u, const float* wdw, const float* bdw, const float* bn_w, const float* bn_b,
                         const float* bn_mean, const float* bn_var, float eps, void* out, int B, int T, int d, int k,
                         void* stream);
int b2c_convmod_dw_fwd(const void* u, const float* wdw, const float* bdw, void* y, double* stats, int B, int T, int d,
                       int k, void* stream);
int b2c_convmod_bn_finalize(const double* stats, int64_t count, float eps, float momentum, float* mean, float* rstd,
                            float* running_mean, float* running_var, int d, void* stream);
int b2c_convmod_bn_apply(const void* y, const float* mean, const float* rstd, const float* bn_w, const float* bn_b,
                         void* out, int64_t rows, int d, void* stream);
/* _bn_finalize + _bn_apply in one launch (training forward): bit-identical results, mean / rstd / running statistics written
 * by the kernel's first CTA */
int b2c_convmod_bn_finalize_apply(const void* y, const double* stats, int64_t count, float eps, float momentum,
                                  const float* bn_w, const float* bn_b, void* out, float* mean, float* rstd,
                                  float* running_mean, float* running_var, int64_t rows, int d, void* stream);
/* backward of the chain.  dc = grad wrt the chain output (bf16 [rows,d]); y = saved depthwise output; mean/rstd = the
 * statistics the forward normalised with (batch stats in train mode, running stats -> rstd = 1/sqrt(var+eps) in eval).
 *  _bn_bwd_reduce: fp64 sums [2*d] (zeroed by the caller): S1 = sum dz, S2 = sum dz*yhat, dz = dc*SiLU'(z)
 *  _dw_bwd       : dy = BN'(dz) on the fly (bn_train selects the batch-stat formula), du bf16 [rows,2d] through the
 *                  depthwise conv and GLU; dwdw [d,k], dbdw, dbn_w, dbn_b fp32 are ACCUMULATED with atomics, and so is
 *                  dbias_u (fp32 [2d], optional): the column sums of du = the bias gradient of the pointwise conv
 *                  that produced u (conformer.py:44-51). */
int b2c_convmod_bn_bwd_reduce(const void* dc, const void* y, const float* mean, const float* rstd, const float* bn_w,
                              const float* bn_b, double* sums, int64_t rows, int d, void* stream);
int b2c_convmod_dw_bwd(const void* dc, const void* y, const void* u, const float* mean, const float* rstd,
                       const float* bn_w, const float* bn_b, const double* sums, int bn_train, const float* wdw,
                       void* du, float* dwdw, float* dbdw, float* dbn_w, float* dbn_b, float* dbias_u, int B, int T, int d,
                       int k, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Multi-head self-attention core, flash style: softmax(Q K^T * scale + key-length mask) V per (batch, head).
 * Replaces scaled_dot_product_attention as reached from nn.MultiheadAttention (torch/nn/functional.py:6608-6682,
 * conformer.py:152,194-200).  Keys t >= lengths[b] are masked; padded QUERY rows are still computed.
 *   qkv  bf16 [B*T, ld]: head h of Q at column h*64, K at k_off + h*64, V at v_off + h*64 (head dim padded to 64)
 *   out  bf16 [B*T, ld_out], head h at column h*64;   lse fp32 [B,H,T] (log2 domain), may be NULL in eval
 *   dropout on the probabilities (train only): p_drop / seed, regenerated identically by the backward kernel.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_attention_fwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths, void* out,
                      int64_t ld_out, float* lse, int B, int T, int H, float scale, float drop_p, uint64_t drop_seed,
                      void* stream);

/* backward: dqkv bf16 [B*T, ld] laid out like qkv (every row < B*T and all three head slabs are written);
 * workspaces: delta fp32 [B,H,T], dq_f32 fp32 [B*T, H*64] (zeroed by this call; every 128-key block adds its partial dQ
 * tile with a TMA reduce, a second kernel casts the sum to bf16).  Recomputes P from lse; the dropout mask is regenerated
 * from (drop_p, drop_seed). */
int b2c_attention_bwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                      const void* out, const void* dout, int64_t ld_out, const float* lse, float* delta, float* dq_f32,
                      void* dqkv, int B, int T, int H, float scale, float drop_p, uint64_t drop_seed, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Relative-positional variant (north_star / SURVEY section 8 row a7).  The reference only ships the sinusoid table
 * (/root/reference/unused_lib/relativepositionalembeddings.py:13-32, used additively at unused_lib/conformer.py:101-103)
 * and imports -- never calls -- the HF module whose arithmetic this follows:
 * Wav2Vec2ConformerSelfAttention._apply_relative_embeddings (transformers modeling_wav2vec2_conformer.py:528-565):
 *     score[i,j] = ( (q_i + u) . k_j  +  BD[i, T-1-i+j] ) * scale,      BD = (q + v) Pw^T,   Pw = linear_pos(PE)
 * Parity for this entry point is pinned to HF (tests/golden/relpos_hf.npz), not to the reference.
 *   qkv  bf16 [B*T, ld]: q+u at column h*64, k at k_off + h*64, v at v_off + h*64, q+v at qv_off + h*64 (the host
 *        side runs the in-projection with augmented weights, so no extra elementwise pass exists)
 *   pw   bf16 [pw_rows >= 2T-1, ld_pw]: row r = projected embedding of relative position (T-1-r), head h at h*64
 * _fwd fuses the BD product into the flash kernel (one extra M128xN256 UMMA per tile, rel_shift as index math).
 * _bwd takes BD materialised by the caller (bd fp32 [B*T, H, Rp], column r, produced with b2c_gemm_bf16 from the same
 * operands) and writes dS at the shifted index into dbd (bf16 [B*T, H, Rp], ZEROED by the caller); the caller then gets
 * d(q+v) = dbd Pw and dPw = dbd^T (q+v) from b2c_gemm_bf16.  dqkv receives d(q+u), dk, dv.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_attention_rel_fwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, int64_t qv_off, const void* pw,
                          int64_t ld_pw, int64_t pw_rows, const int32_t* lengths, void* out, int64_t ld_out, float* lse,
                          int B, int T, int H, float scale, float drop_p, uint64_t drop_seed, void* stream);
int b2c_attention_rel_bwd(const void* qkv, int64_t ld, int64_t k_off, int64_t v_off, const int32_t* lengths,
                          const void* out, const void* dout, int64_t ld_out, const float* lse, float* delta,
                          float* dq_f32, void* dqkv, const float* bd, void* dbd, int64_t Rp, int B, int T, int H,
                          float scale, float drop_p, uint64_t drop_seed, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * ConvSubSampling (/root/reference/lib/convsubsampling.py:16-45): conv_sub_2(conv_sub_1(x)), no activation
 * between, valid padding.  Executed as the exact composition Conv2d(1->C2, kc x kc, stride s1*s2) with
 * kc = k1 + (k2-1)*s1: _compose builds Wc (bf16 [C2, KP], KP = kc*kc rounded up to 64)
 * and bc (fp32 [C2]) from the module's own parameters; _fwd is an implicit GEMM on tcgen05.
 *   x   fp32 [B, F, T]         (in_size = 1)
 *   out [B, T', F', C2] token-major (row (b*T'+t')*F'+f', channels contiguous), bf16 (out_f32=0) or fp32 (1);
 *       the reference module's (B, C2, F', T') return value is the permuted VIEW of the fp32 form.
 * _compose_bwd maps dWc/dbc (fp32 [C2,KP] / [C2]) back to dW1, db1, dW2, db2 (accumulating); _wgrad accumulates dWc.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_convsub_compose(const float* W1, const float* b1, const float* W2, const float* b2, int C1, int C2, int k1,
                        int s1, int k2, void* Wc, float* Wc_f32, float* bc, int KP, void* stream);
int b2c_convsub_fwd(const float* x, const void* Wc, const float* bc, void* out, int out_f32, int B, int F, int T,
                    int C2, int kc, int stride, int KP, void* stream);
/* dWc fp32 [C2,KP] += dout^T * im2col(x); dout bf16 token-major [B*T', F'*C2]. (dbc = column sums: b2c_colsum_bf16) */
int b2c_convsub_wgrad(const float* x, const void* dout, float* dWc, int B, int F, int T, int C2, int kc, int stride,
                      int KP, void* stream);
int b2c_convsub_compose_bwd(const float* dWc, const float* dbc, const float* W1, const float* b1, const float* W2,
                            int C1, int C2, int k1, int s1, int k2, int KP, float* dW1, float* db1, float* dW2,
                            float* db2, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * SpecAugment (/root/reference/lib/standard/asrnn.py:91-192): time-warp gather + frequency / time masks in one
 * pass.  The random draws stay on the host (CPython `random` stream, bit-exact: b2c_mt19937_*); the kernel gets
 * the drawn parameters.
 *   x, out fp32 [B,F,T];  tau int32 [B]
 *   warp_w, warp_w0 int32 [n_warp, B]  (w, w0 of asrnn.py:103-107), composed in call order
 *   fmask int32 [n_fmask, 2] = (f0, f);  tmask int32 [n_tmask, B, 2] = (t0, t)
 *   apply_masks 0 = "reference" mode (the reference's masks are no-ops), 1 = "paper" mode
 * ------------------------------------------------------------------------------------------------------- */
int b2c_specaug_apply(const float* x, float* out, const int32_t* tau, const int32_t* warp_w, const int32_t* warp_w0,
                      int n_warp, const int32_t* fmask, int n_fmask, const int32_t* tmask, int n_tmask,
                      int apply_masks, float mask_value, int B, int F, int T, void* stream);

/* CPython-compatible MT19937 (host): seed(int) via init_by_array, randint via the bit_length rejection loop
 * (/usr/lib/python3.12/random.py:242-249,318,336).  state_host: 625 uint32 words owned by the caller. */
void b2c_mt19937_seed_host(uint32_t* state_host, uint64_t seed);
int64_t b2c_mt19937_randint_host(uint32_t* state_host, int64_t a, int64_t b);

/* ---------------------------------------------------------------------------------------------------------
 * Fused multi-tensor Adafactor step -- transformers.Adafactor(lr, beta1, scale_parameter=False, relative_step=False)
 * as the reference constructs it (/root/reference/lib/standard/runner.py:36,146).  Four launches for the whole
 * parameter set.  tensors_dev: device array of 96-byte records
 *   { float* p; const float* g; float* exp_avg; float* row_or_sq; float* col; float* rowsum_ws; float* colsum_ws;
 *     float* rmean_ws; float* sumsq_ws; int32 P, R, C, factored; int64 n }      (tensor viewed as [P][R][C])
 * chunks_dev: device array of { int32 tensor, row0, nrows, pad } covering every flattened row once.  The workspace
 * fields must be zero on entry.  step_dev (int64, starts at 0) is incremented on the device and beta2t_dev receives
 * 1 - step^decay_rate, so a captured CUDA graph advances the schedule on every replay.  Five launches.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_adafactor_step(const void* tensors_dev, int n_tensors, const void* chunks_dev, int n_chunks, int64_t* step_dev,
                       float* beta2t_dev, float lr, float beta1, float eps1, float clip, float decay_rate, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * One call per Conformer layer, eval mode (torchaudio ConformerLayer.forward, conformer.py:176-212, dropout inactive,
 * BatchNorm on running statistics): enqueues the layer's kernels on `stream`.  For the eager inference / NST
 * pseudo-labelling path (Runner.generate_labels, runner.py:255-281), where per-launch host overhead dominates.
 *   x        fp32 [B*T, d] residual stream, updated IN PLACE (layer input -> layer output)
 *   weights  GEMM weights bf16 row-major (out, in), attention heads zero-padded to 64 (in_w [3*H*64, d], in_b [3*H*64],
 *            out_w [d, H*64]); LayerNorm / bias / depthwise / BatchNorm tensors fp32 as stored by the module
 *   workspace >= b2c_conformer_layer_eval_workspace(...) bytes, 256-byte aligned, reusable across layers
 * ------------------------------------------------------------------------------------------------------- */
typedef struct b2c_ffn_weights {
  const float* ln_w; const float* ln_b;
  const void* w1; const float* b1;          /* Linear(d, f) */
  const void* w2; const float* b2;          /* Linear(f, d) */
} b2c_ffn_weights;
typedef struct b2c_layer_weights {
  b2c_ffn_weights ffn1;
  const float* att_ln_w; const float* att_ln_b;
  const void* in_w; const float* in_b; const void* out_w; const float* out_b;
  const float* conv_ln_w; const float* conv_ln_b;
  const void* pw1_w; const float* pw1_b;    /* Conv1d(d, 2d, 1) */
  const float* dw_w; const float* dw_b;     /* depthwise Conv1d(d, d, k): [d, k] */
  const float* bn_w; const float* bn_b; const float* bn_mean; const float* bn_var;
  const void* pw2_w; const float* pw2_b;    /* Conv1d(d, d, 1) */
  b2c_ffn_weights ffn2;
  const float* final_ln_w; const float* final_ln_b;
} b2c_layer_weights;
size_t b2c_conformer_layer_eval_workspace(int B, int T, int d, int f, int H);
int b2c_conformer_layer_eval(float* x, const b2c_layer_weights* weights, const int32_t* lengths, int B, int T, int d, int f,
                             int H, int head_dim, int k, float bn_eps, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Small memory-bound helpers
 * ------------------------------------------------------------------------------------------------------- */
/* multi-tensor fp32 -> bf16 cast; optional zero padding of rows/cols (pads attention heads to 64):
 *   dst[r_dst(r), c] with r in [0,rows), src row pitch = cols, dst pitch = dst_cols >= cols */
int b2c_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
/* one launch for a whole parameter set: table_dev = device array of {const float* src; bf16* dst; int64 n} (24 B each,
 * sources/destinations 16-byte aligned); max_n = largest n (sizes the grid) */
int b2c_multi_cast_f32_bf16(const void* table_dev, int n_tensors, int64_t max_n, void* stream);
/* dst bf16 [rows, cols] = alpha * dropout(src fp32 [rows, cols]) -- gradient of  x + alpha*drop(y)  wrt y;
 * colsum (optional fp32 [cols]) += column sums of dst: the bias gradient of the linear layer that produced y */
int b2c_scale_cast(const float* src, void* dst, int64_t rows, int64_t cols, float alpha, float drop_p,
                   uint64_t drop_seed, float* colsum, void* stream);
/* out = x + e (fp32, n % 4 == 0): `x = x + rel_pos_emb` of the reference's own positional variant
 * (unused_lib/conformer.py:101-103, table from unused_lib/relativepositionalembeddings.py:26-32) */
int b2c_add_f32(const float* x, const float* e, float* out, int64_t n, void* stream);
/* out fp32 [cols] += column sums of a bf16 [rows, ld] matrix (bias gradients) */
int b2c_colsum_bf16(const void* a, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream);
/* pads / unpads per-head columns:  dst[r, h*dst_dh + c] = src[r, h*src_dh + c] (zero fill), nblk = blocks per row */
int b2c_repack_heads(const void* src, void* dst, int64_t rows, int nblk, int src_dh, int dst_dh, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused log-softmax + CTC loss (+ gradient w.r.t. the logits): `nn.LogSoftmax(-1)` at asrnn.py:44,254-256 followed by
 * `CTCLoss(blank=hp.blank_idx, zero_infinity=True)` at runner.py:35,143 (reduction 'mean': nll_b / max(S_b, 1), averaged
 * over the batch; infinite losses and their gradients zeroed when zero_infinity).
 *   logits fp32 [B, T, V] batch-major (row = b*T + t); log-softmaxed input gives the same result (log_softmax is idempotent)
 *   targets int32 [B, Smax] (labels beyond target_lens[b] ignored), input_lens / target_lens int32 [B]
 *   loss fp32 [1], nll fp32 [B] (optional), grad_logits fp32 [B, T, V] = d loss / d logits * (*grad_out)
 *   phases: 1 forward only (state posteriors stay in `workspace`), 2 gradient only (same workspace), 3 both
 *   grad_out: DEVICE scalar or NULL (= 1).  workspace: b2c_ctc_workspace(B, T, Smax) bytes.
 * ------------------------------------------------------------------------------------------------------- */
size_t b2c_ctc_workspace(int B, int T, int Smax);
int b2c_ctc_loss(const float* logits, const int32_t* targets, const int32_t* input_lens, const int32_t* target_lens, int B,
                 int T, int V, int Smax, int blank, int zero_infinity, int phases, const float* grad_out, float* loss,
                 float* nll, float* grad_logits, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm1d over token rows (+ the SiLU in front of it): tail of the reference's projection block
 * `projection_fc -> swish -> projection_batch_norm` (asrnn.py:73-89,217).  Linear + SiLU = b2c_gemm_bf16 with act = SILU and
 * out2 = z (saved pre-activation); h bf16 [rows, C] is its output.
 *   _stats       stats fp64 [2C] (zeroed by the caller) += per-channel sum, sum of squares of h; b2c_convmod_bn_finalize turns
 *                them into mean / rstd and updates the running statistics
 *   _apply       y fp32 [rows, C] = (h - mean) * rstd * gamma + beta
 *   _bwd_reduce  sums fp64 [2C] (zeroed) += sum dy (= dbeta), sum dy * hhat (= dgamma)
 *   _bwd_apply   dz bf16 = BN'(dy) * SiLU'(z)  (train = 1: batch-statistics formula, 0: gamma * rstd * dy);
 *                colsum fp32 [C] (optional) += column sums of dz (the Linear's bias gradient)
 * ------------------------------------------------------------------------------------------------------- */
int b2c_bn1d_stats(const void* h, double* stats, int64_t rows, int C, void* stream);
int b2c_bn1d_apply(const void* h, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y,
                   int64_t rows, int C, void* stream);
int b2c_bn1d_bwd_reduce(const float* dy, const void* h, const float* mean, const float* rstd, double* sums, int64_t rows,
                        int C, void* stream);
int b2c_bn1d_bwd_apply(const float* dy, const void* h, const void* z, const float* mean, const float* rstd,
                       const float* gamma, const double* sums, int train, void* dz, float* colsum, int64_t rows, int C,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LSTM cell (the bi-LSTM "decoder", self.lstm = nn.LSTM(...) at /root/reference/lib/standard/asrnn.py:38,
 * called at asrnn.py:252).  One call per time step and direction, between two recurrent GEMMs (b2c_gemm_bf16); torch gate
 * order i, f, g, o along the 4H axis.  N rows (the LSTM's batch axis), H hidden units (H % 4 == 0).
 *   _fwd  gates fp32 [N, 4H] = x_t W_ih^T + b_ih + b_hh + h_{t-1} W_hh^T;  c_prev fp32 [N, H] or NULL (= 0).
 *         c_out fp32 [N, H]; h_out fp32, row pitch ld_h (the [L, N, dirs*H] output, this direction's column block);
 *         h_bf16 bf16 [N, H] (operand of the next step's GEMM); acts fp32 [N, 4H] or NULL: activated gates, saved for _bwd
 *   _bwd  dout fp32 (row pitch ld_dout) or NULL: gradient of this step's output; dh_rec fp32 [N, H] or NULL: dG_{t+1} W_hh;
 *         dc fp32 [N, H]: IN d loss / d c_t carried from step t+1 (zero at the last step), OUT d loss / d c_{t-1};
 *         acts / c_prev (NULL = 0) / c_cur as saved by _fwd; dgates bf16 [N, 4H]: gradient of the pre-activations
 * ------------------------------------------------------------------------------------------------------- */
int b2c_lstm_cell_fwd(const float* gates, const float* c_prev, float* c_out, float* h_out, int64_t ld_h, void* h_bf16,
                      float* acts, int64_t N, int H, void* stream);
int b2c_lstm_cell_bwd(const float* dout, int64_t ld_dout, const float* dh_rec, float* dc, const float* acts,
                      const float* c_prev, const float* c_cur, void* dgates, int64_t N, int H, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Input pipeline.
 * b2c_logmel: librosa.feature.melspectrogram -> log -> min-max normalisation of `SpeechCommands.get_mels`
 *   (speechcommands.py:113-119) for n packed waveforms (wave fp32, wave_off int64 [n+1] sample offsets), written into the
 *   zero-padded batch out fp32 [B, n_mels, t_max] (rows >= n zero; the padding of get_batch, speechcommands.py:186-190);
 *   frames_out int32 [B].  basis fp32 [n_mels, n_fft/2+1] with per-filter non-zero ranges f_lo/f_hi int32 [n_mels];
 *   mel_scratch fp32 [n, n_mels, frames_max], frames_max >= 1 + max_len // hop.  n_fft a power of two <= 4096.
 * b2c_pad_batch: packed ragged features (utterance i = [F, lens[i]] at off[i]) -> out fp32 [B, F, t_max] zero-padded.
 * ------------------------------------------------------------------------------------------------------- */
int b2c_logmel(const float* wave, const int64_t* wave_off, int n, int B, const float* basis, const int32_t* f_lo,
               const int32_t* f_hi, int n_fft, int hop, int n_mels, int frames_max, int t_max, float* mel_scratch, float* out,
               int32_t* frames_out, void* stream);
int b2c_pad_batch(const float* packed, const int64_t* off, const int32_t* lens, int n, int B, int F, int t_max, float* out,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CONFORMER_H_ */
