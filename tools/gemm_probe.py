"""Tiny driver for ncu: a few launches of the dominant GEMM shapes at Conformer-M sizes."""
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
dW = torch.zeros(d, f, device="cuda")
for _ in range(3):
    ops.gemm(a, w, M=M, N=f, K=d, bias=bias, act=ops.ACT_SILU, out=out, block_n=128)            # FFN-1
    ops.gemm(a, w, M=M, N=f, K=d, bias=bias, act=ops.ACT_SILU, out=out, block_n=256)            # FFN-1, wide tile
    ops.gemm(h, w2, M=M, N=d, K=f, bias=None, resid=x, out=x, alpha=0.5)                        # FFN-2
    ops.gemm(a, w2, M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=h, out=out)             # dgrad through SiLU
    ops.gemm(a, h, M=d, N=f, K=M, a_mn=True, b_mn=True, out=dW, accumulate=True, split_k=18)    # wgrad
torch.cuda.synchronize()
print("ok")
