"""cuBLAS (torch.matmul, bf16) at the Conformer-M GEMM shapes of the training step -- a YARDSTICK for the hand-written
tcgen05 GEMM (never on the product path).  Same timing recipe as tools/kernel_bench.py (graph-replayed back-to-back launches,
L2 flushed).  Also times a chain of N tiny dependent kernels in one graph: the per-boundary launch bubble."""
import sys
import torch
sys.path.insert(0, ".")
from tools.kernel_bench import timeit  # noqa: E402

dev, bf = "cuda", torch.bfloat16
M = 23872
shapes = [("ffn1/fwd   N1024 K256", M, 1024, 256), ("ffn2/fwd   N256 K1024", M, 256, 1024), ("qkv/fwd    N768 K256", M, 768, 256),
          ("proj/fwd   N256 K256", M, 256, 256), ("pw1/fwd    N512 K256", M, 512, 256), ("wgrad 256x1024 K=M", 256, 1024, M),
          ("wgrad 1024x256 K=M", 1024, 256, M), ("wgrad 256x256 K=M", 256, 256, M), ("bridge N256 K2304", M, 256, 2304)]
for name, m, n, k in shapes:
    a = torch.randn(m, k, device=dev).to(bf)
    b = torch.randn(n, k, device=dev).to(bf)
    out = torch.empty(m, n, device=dev, dtype=bf)
    us = timeit(lambda: torch.matmul(a, b.t(), out=out))
    print(f"cublas {name:24s} {us:8.2f} us  {2.0 * m * n * k / us / 1e6:8.1f} TF/s", flush=True)
# launch bubble: 200 dependent tiny kernels
x = torch.zeros(1024, device=dev)
us = timeit(lambda: x.add_(1.0), reps=200)
print(f"tiny dependent kernel in a graph: {us:.2f} us per launch (kernel + boundary)")
