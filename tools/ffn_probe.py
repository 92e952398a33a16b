"""Scaling probe of b2c_ffn_fused_eval: time vs number of token tiles and hidden chunks (graph-replayed, L2-flushed)."""
import math, sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from tools.kernel_bench import timeit
bf = torch.bfloat16
for M, d, f in [(128, 256, 128), (128, 256, 1024), (128 * 148, 256, 128), (128 * 148, 256, 1024), (23872, 256, 1024), (128 * 148, 256, 2048), (23872, 144, 576)]:
    x = torch.randn(M, d, device="cuda")
    ln_w, ln_b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    w1, b1 = (torch.randn(f, d, device="cuda") / math.sqrt(d)).to(bf), torch.zeros(f, device="cuda")
    w2, b2 = (torch.randn(d, f, device="cuda") / math.sqrt(f)).to(bf), torch.zeros(d, device="cuda")
    out = torch.empty_like(x)
    us = timeit(lambda: ops.ffn_fused_eval(x, ln_w, ln_b, w1, b1, w2, b2, out=out))
    a = torch.empty(M, d, device="cuda", dtype=bf); h = torch.empty(M, f, device="cuda", dtype=bf)
    def unf():
        ops.layernorm_fwd(x, ln_w, ln_b, out=a)
        ops.gemm(a, w1, M=M, N=f, K=d, bias=b1, act=ops.ACT_SILU, out=h)
        ops.gemm(h, w2, M=M, N=d, K=f, bias=b2, resid=x, out=out, alpha=0.5)
    us2 = timeit(unf)
    print(f"M={M:6d} d={d} f={f:5d}: fused {us:8.2f} us  ({4.0*M*d*f/us/1e6:7.1f} TF/s)   un-fused {us2:8.2f} us", flush=True)
