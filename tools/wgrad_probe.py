"""Weight-gradient GEMM tile / split-K sweep at the Conformer-M shapes (out[Mo,No] += dY[tokens,Mo]^T X[tokens,No])."""
import sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from tools.kernel_bench import timeit
M = 23872
bf = torch.bfloat16
for Mo, No in [(256, 256), (768, 256), (512, 256), (256, 1024), (1024, 256)]:
    dy = torch.randn(M, Mo, device="cuda").to(bf); x = torch.randn(M, No, device="cuda").to(bf)
    out = torch.zeros(Mo, No, device="cuda")
    best = None
    for bn in (-256, -128, 128, 256, 64):
        for split in (9, 18, 37, 74, 93):
            try:
                us = timeit(lambda: ops.gemm(dy, x, M=Mo, N=No, K=M, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=split, block_n=bn))
            except Exception as e:
                continue
            tf = 2 * M * Mo * No / us / 1e6
            print(f"wgrad {Mo}x{No}  bn{bn:5d} split{split:3d}  {us:7.2f} us  {tf:6.1f} TF/s", flush=True)
            if best is None or us < best[0]:
                best = (us, bn, split)
    print(f"== best {Mo}x{No}: {best}")
