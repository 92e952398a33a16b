"""Weight-stationary GEMM path (block_n = 9000 + BN) against the streaming path (block_n = BN) at the Conformer-M shapes with
K = 256: checks the results agree with each other and with torch, then times both (one CUDA graph of back-to-back launches,
L2 flushed before each timed replay).  Usage on the GPU box: python tools/bstat_probe.py"""
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
    d, f = 256, 1024
    for M in (1000, 23872):
        a_d = (torch.randn(M, d, device=dev) / 4).to(bf)
        pre = torch.randn(M, f, device=dev).to(bf)
        x = torch.randn(M, d, device=dev)
        w = {n: (torch.randn(n, d, device=dev) / 16).to(bf) for n in (256, 512, 768, 1024)}
        wt = {n: (torch.randn(d, n, device=dev) / 16).to(bf) for n in (256, 1024)}        # [K, N] storage (dgrad)
        bias = {n: torch.randn(n, device=dev) for n in (256, 512, 768, 1024)}
        cases = []

        def plain(N):
            def run(bn):
                out = torch.empty(M, N, device=dev, dtype=bf)
                ops.gemm(a_d, w[N], M=M, N=N, K=d, bias=bias[N], out=out, block_n=bn)
                return (out,)
            return run

        def silu(N, train):
            def run(bn):
                out = torch.empty(M, N, device=dev, dtype=bf)
                out2 = torch.empty(M, N, device=dev, dtype=bf) if train else None
                ops.gemm(a_d, w[N], M=M, N=N, K=d, bias=bias[N], act=ops.ACT_SILU, out=out, out2=out2,
                         drop_p=0.5 if train else 0.0, drop_seed=11, block_n=bn)
                return (out, out2) if train else (out,)
            return run

        def resid(N, train):
            def run(bn):
                out = torch.empty(M, N, device=dev)
                ops.gemm(a_d, w[N], M=M, N=N, K=d, bias=bias[N], resid=x, out=out, alpha=0.5, drop_p=0.5 if train else 0.0,
                         drop_seed=5, block_n=bn)
                return (out,)
            return run

        def dgrad_plain(N):
            def run(bn):
                out = torch.empty(M, N, device=dev, dtype=bf)
                ops.gemm(a_d, wt[N], M=M, N=N, K=d, b_mn=True, out=out, block_n=bn)
                return (out,)
            return run

        def dgrad_silu(N):
            def run(bn):
                out = torch.empty(M, N, device=dev, dtype=bf)
                cs = torch.zeros(N, device=dev)
                ops.gemm(a_d, wt[N], M=M, N=N, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=pre, out=out, alpha=2.0, colsum=cs,
                         block_n=bn)
                return (out, cs)
            return run

        cases = [("ffn1 silu eval   N1024", silu(1024, False), 1024), ("ffn1 silu train  N1024", silu(1024, True), 1024),
                 ("qkv plain        N768 ", plain(768), 768), ("pw1 plain        N512 ", plain(512), 512),
                 ("proj resid eval  N256 ", resid(256, False), 256), ("proj resid train N256 ", resid(256, True), 256),
                 ("dgrad plain      N256 ", dgrad_plain(256), 256), ("dgrad silu_bwd   N1024", dgrad_silu(1024), 1024)]
        for name, run, N in cases:
            ref = run(256 if N >= 512 else 128)
            line = f"M{M:<6d} {name}"
            for bs in (9128, 9256):
                try:
                    got = run(bs)
                except ValueError:                              # no stationary kernel for this epilogue (fp32 residual outputs)
                    line += f"  bn{bs} n/a            "
                    continue
                err = 0.0
                for r, g in zip(ref, got):
                    if r is None:
                        continue
                    r32, g32 = r.float(), g.float()
                    fin = torch.isfinite(r32) & (r32.abs() < 1e3)       # dropped pre-activations carry a -1e4 marker
                    same_marker = bool(((r32.abs() >= 1e3) == (g32.abs() >= 1e3)).all())
                    e = float((r32 - g32)[fin].abs().max() / (r32[fin].abs().max() + 1e-9)) if same_marker else float("inf")
                    err = max(err, e)
                line += f"  bn{bs} rel-err {err:.1e}"
            if M > 10000:
                t0 = timeit(lambda: run(256 if N >= 512 else 128))
                try:
                    t1 = timeit(lambda: run(9128))
                    t2 = timeit(lambda: run(9256))
                except ValueError:
                    t1 = t2 = float("nan")
                fl = 2 * M * N * d
                line += f"   stream {t0:6.2f} us ({fl / t0 / 1e6:5.0f} TF/s)  stat128 {t1:6.2f} us ({fl / t1 / 1e6:5.0f})  stat256 {t2:6.2f} us ({fl / t2 / 1e6:5.0f})"
            print(line, flush=True)
    print("OK")


if __name__ == "__main__":
    main()
