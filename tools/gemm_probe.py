"""Tiny driver for ncu: a few launches of the dominant kernels at Conformer-M shapes."""
import sys
import torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
M, d, f = 23872, 256, 1024
bf = torch.bfloat16
a = torch.randn(M, d, device="cuda").to(bf)
w = torch.randn(f, d, device="cuda").to(bf)
bias = torch.randn(f, device="cuda")
out = torch.empty(M, f, device="cuda", dtype=bf)
h = torch.randn(M, f, device="cuda").to(bf)
w2 = torch.randn(d, f, device="cuda").to(bf)
x = torch.randn(M, d, device="cuda")
for _ in range(4):
    ops.gemm(a, w, M=M, N=f, K=d, bias=bias, act=ops.ACT_SILU, out=out)            # FFN-1
    ops.gemm(h, w2, M=M, N=d, K=f, bias=None, resid=x, out=x, alpha=0.5)            # FFN-2
torch.cuda.synchronize()
print("ok")
