import sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from tools.kernel_bench import timeit
M, d, f = 23872, 256, 1024
bf = torch.bfloat16; dev = "cuda"
a_d = torch.randn(M, d, device=dev).to(bf); w_df = torch.randn(d, f, device=dev).to(bf)
o_f = torch.empty(M, f, device=dev, dtype=bf); o2_f = torch.randn(M, f, device=dev).to(bf); cs_f = torch.zeros(f, device=dev)
for name, kw in [("plain", {}), ("drop", dict(drop_p=0.5, drop_seed=1)), ("colsum", dict(colsum=cs_f)), ("drop+colsum", dict(drop_p=0.5, drop_seed=1, colsum=cs_f))]:
    us = timeit(lambda: ops.gemm(a_d, w_df, M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=o2_f, out=o_f, **kw))
    print(f"du dgrad {name:12s} {us:6.2f} us")
