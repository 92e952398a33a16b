import sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from tools.kernel_bench import timeit
M, d, f = 23872, 256, 1024
bf = torch.bfloat16; dev = "cuda"
a = torch.randn(M, d, device=dev).to(bf); w = torch.randn(f, d, device=dev).to(bf); b = torch.randn(f, device=dev)
o = torch.empty(M, f, device=dev, dtype=bf); o2 = torch.empty(M, f, device=dev, dtype=bf)
for name, kw in [("silu", {}), ("silu+out2", dict(out2=o2)), ("silu+drop", dict(drop_p=0.5, drop_seed=1)),
                 ("silu+out2+drop", dict(out2=o2, drop_p=0.5, drop_seed=1)), ("bias only", dict(act=ops.ACT_NONE))]:
    kw2 = dict(act=ops.ACT_SILU); kw2.update(kw)
    for bn in (128, 256):
        us = timeit(lambda: ops.gemm(a, w, M=M, N=f, K=d, bias=b, out=o, block_n=bn, **kw2))
        print(f"ffn1 {name:16s} bn{bn}: {us:6.2f} us", flush=True)
