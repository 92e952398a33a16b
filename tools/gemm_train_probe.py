"""Tile-width sweep of the forward / dgrad GEMMs WITH their training epilogues (dropout, saved pre-activation, fused
column sums) at the Conformer-M shapes."""
import sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from tools.kernel_bench import timeit
M, d, f = 23872, 256, 1024
bf = torch.bfloat16
dev = "cuda"
a_d = torch.randn(M, d, device=dev).to(bf); a_f = torch.randn(M, f, device=dev).to(bf)
a_3d = torch.randn(M, 3 * d, device=dev).to(bf); a_2d = torch.randn(M, 2 * d, device=dev).to(bf)
w_fd = torch.randn(f, d, device=dev).to(bf); w_df = torch.randn(d, f, device=dev).to(bf); w_dd = torch.randn(d, d, device=dev).to(bf)
w_3d = torch.randn(3 * d, d, device=dev).to(bf); w_2d = torch.randn(2 * d, d, device=dev).to(bf)
b_f, b_d = torch.randn(f, device=dev), torch.randn(d, device=dev)
x = torch.randn(M, d, device=dev); xo = torch.empty(M, d, device=dev)
o_f = torch.empty(M, f, device=dev, dtype=bf); o2_f = torch.empty(M, f, device=dev, dtype=bf); o_d = torch.empty(M, d, device=dev, dtype=bf)
o_3d = torch.empty(M, 3 * d, device=dev, dtype=bf); o_2d = torch.empty(M, 2 * d, device=dev, dtype=bf)
cs_f, cs_d = torch.zeros(f, device=dev), torch.zeros(d, device=dev)
cases = {
    "ffn1 fwd  silu+out2+drop": lambda bn: ops.gemm(a_d, w_fd, M=M, N=f, K=d, bias=b_f, act=ops.ACT_SILU, out=o_f, out2=o2_f, drop_p=0.5, drop_seed=1, block_n=bn),
    "ffn2 fwd  resid+drop    ": lambda bn: ops.gemm(a_f, w_df, M=M, N=d, K=f, bias=b_d, resid=x, out=xo, alpha=0.5, drop_p=0.5, drop_seed=2, block_n=bn),
    "qkv  fwd  bias          ": lambda bn: ops.gemm(a_d, w_3d, M=M, N=3 * d, K=d, bias=b_f[:3 * d], out=o_3d, block_n=bn),
    "pw1  fwd  bias          ": lambda bn: ops.gemm(a_d, w_2d, M=M, N=2 * d, K=d, bias=b_f[:2 * d], out=o_2d, block_n=bn),
    "proj fwd  resid+drop    ": lambda bn: ops.gemm(a_d, w_dd, M=M, N=d, K=d, bias=b_d, resid=x, out=xo, drop_p=0.5, drop_seed=3, block_n=bn),
    "du   dgrad silu'+drop+cs": lambda bn: ops.gemm(a_d, w_df, M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=o2_f, out=o_f, drop_p=0.5, drop_seed=1, colsum=cs_f, block_n=bn),
    "da   dgrad K=1024       ": lambda bn: ops.gemm(a_f, w_fd, M=M, N=d, K=f, b_mn=True, out=o_d, block_n=bn),
    "da   dgrad K=768 (qkv)  ": lambda bn: ops.gemm(a_3d, w_3d, M=M, N=d, K=3 * d, b_mn=True, out=o_d, block_n=bn),
    "da   dgrad K=512 (pw1)  ": lambda bn: ops.gemm(a_2d, w_2d, M=M, N=d, K=2 * d, b_mn=True, out=o_d, block_n=bn),
    "dc   dgrad K=256        ": lambda bn: ops.gemm(a_d, w_dd, M=M, N=d, K=d, b_mn=True, out=o_d, block_n=bn),
}
for name, fn in cases.items():
    res = []
    for bn in (0, 64, 128, 256):
        try:
            res.append((bn, timeit(lambda: fn(bn))))
        except Exception as e:
            res.append((bn, float("nan")))
    print(name, "  ".join(f"bn{bn}: {us:6.2f} us" for bn, us in res), flush=True)
