"""Does sharing the weight tile between CTAs pay on the deep-K shapes?  Times the streaming tiles (block_n 128 / 256), CTA
pairs (-256: cta_group::2, each CTA loads half of B) and multicast clusters (2256 / 4256) on the Conformer-M GEMMs with
K >= 512 and on the weight gradients (one CUDA graph of back-to-back launches, L2 flushed before each replay).
Usage on the GPU box: python tools/share_probe.py"""
import sys

import torch

sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps * 1e3)
    return min(ts)


def main():
    torch.manual_seed(0)
    bf = torch.bfloat16
    M, d, f = 23872, 256, 1024
    h = torch.randn(M, f, device=dev).to(bf)
    xd = torch.randn(M, d, device=dev).to(bf)
    x = torch.randn(M, d, device=dev)
    w2 = (torch.randn(d, f, device=dev) / 32).to(bf)          # FFN2 weight (out=d, in=f)
    w1 = (torch.randn(f, d, device=dev) / 16).to(bf)          # FFN1 weight (out=f, in=d): its dgrad reads it [K=f, N=d]
    b2 = torch.randn(d, device=dev)

    def ffn2(train):
        def run(bn):
            ops.gemm(h, w2, M=M, N=d, K=f, bias=b2, resid=x, out=x, alpha=0.5, drop_p=0.5 if train else 0.0, drop_seed=3, block_n=bn)
        return run

    def dgrad_ffn1(bn):
        out = torch.empty(M, d, device=dev, dtype=bf)
        ops.gemm(h, w1, M=M, N=d, K=f, b_mn=True, out=out, block_n=bn)

    for name, run, fl in (("ffn2 resid eval  N256 K1024", ffn2(False), 2 * M * d * f), ("ffn2 resid train N256 K1024", ffn2(True), 2 * M * d * f),
                          ("dgrad plain      N256 K1024", dgrad_ffn1, 2 * M * d * f)):
        line = name
        for bn in (256, 128, -256, -128, 2256, 4256, 2128):
            try:
                t = timeit(lambda: run(bn))
                line += f"  bn{bn}: {t:6.2f} us ({fl / t / 1e6:4.0f})"
            except Exception as e:  # noqa: BLE001
                line += f"  bn{bn}: n/a"
        print(line, flush=True)
    # weight gradients dW[Nw, Kw] += dY^T X over the M tokens
    for Nw, Kw in ((1024, 256), (256, 1024), (768, 256), (256, 256), (512, 256)):
        dy = torch.randn(M, Nw, device=dev).to(bf)
        xx = torch.randn(M, Kw, device=dev).to(bf)
        out = torch.zeros(Nw, Kw, device=dev)
        fl = 2 * M * Nw * Kw
        line = f"wgrad {Nw:4d} x {Kw:4d}"
        kb = (M + 63) // 64
        for bn in (0, 256, -256, 128):
            tile_n = 128 if bn == 128 else 256
            tiles = ((Nw + (255 if bn == -256 else 127)) // (256 if bn == -256 else 128)) * ((Kw + tile_n - 1) // tile_n)
            sk = max(1, min((74 if bn == -256 else 148) // tiles, kb // 4))
            try:
                kw = dict(block_n=bn)
                t = timeit(lambda: ops.gemm(dy, xx, M=Nw, N=Kw, K=M, a_mn=True, b_mn=True, out=out, out_dtype=torch.float32, accumulate=True,
                                            split_k=sk, **kw))
                line += f"  bn{bn}: {t:6.2f} us ({fl / t / 1e6:4.0f})"
            except Exception as e:  # noqa: BLE001
                line += f"  bn{bn}: n/a ({str(e)[:40]})"
        print(line, flush=True)
    print("OK")


if __name__ == "__main__":
    main()
