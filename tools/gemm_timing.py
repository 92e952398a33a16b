"""Debug aid (needs a library built with `make EXTRA=-DB2C_GEMM_TIMING`): per-role cycle stamps of CTA 0 for one GEMM.
usage: python tools/gemm_timing.py [block_n] [shape]   shape in {ffn1, ffn2, qkv}"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops, _lib
import os
_lib.LIB_PATH = _lib.LIB_PATH.replace("libb200conformer.so", os.environ.get("B2C_DBG_LIB", "libb200conformer_dbg.so"))
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = sys.argv[2] if len(sys.argv) > 2 else "ffn1"
M, d, f = 23872, 256, 1024
bf = torch.bfloat16
N, K = {"ffn1": (f, d), "ffn1t": (f, d), "ffn2": (d, f), "qkv": (3 * d, d), "dsilu": (f, d), "dgrad": (d, f), "proj": (d, d)}[shape]
a = torch.randn(M, K, device="cuda").to(bf); w = torch.randn(N, K, device="cuda").to(bf)
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=bf)
x = torch.randn(M, N, device="cuda")
wt = torch.randn(K, N, device="cuda").to(bf); aux = torch.randn(M, N, device="cuda").to(bf)
def run():
    if shape == "ffn1": ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out=out, block_n=bn)
    elif shape == "ffn1t": ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out=out, out2=aux, drop_p=0.5, drop_seed=11, block_n=bn)
    elif shape in ("ffn2", "proj"): ops.gemm(a, w, M=M, N=N, K=K, bias=bias, resid=x, out=x, alpha=0.5, block_n=bn)
    elif shape == "dsilu": ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=aux, out=out, block_n=bn)
    elif shape == "dgrad": ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True, out=out, block_n=bn)
    else: ops.gemm(a, w, M=M, N=N, K=K, out=out, block_n=bn)
for _ in range(3): run()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 16 * 32))()
lib = _lib.load()
lib.b2c_debug_gemm_timing.argtypes = [ctypes.c_void_p]
lib.b2c_debug_gemm_timing(buf)
import numpy as np
t = np.array(buf[:], dtype=np.int64).reshape(3, 16, 32)
t0 = t[t > 0].min()
kbs = (K + 63) // 64
print(f"shape {shape} M{M} N{N} K{K} bn{bn}: cycles relative to CTA 0's first stamp")
for it in range(4):
    pr = t[0, it, :2 * kbs] - t0
    mm = t[1, it, :2 + 2 * kbs] - t0
    ep = t[2, it, :12] - t0
    pf = t[2, it, 8:12] - t0
    print(f"tile {it}: producer (wait_empty, got) per kb: {pr.tolist()}")
    print(f"         mma tempty wait {mm[0]}->{mm[1]}; full (wait, got) per kb: {mm[2:].tolist()}")
    print(f"         epilogue warp0 tfull wait {ep[0]}->{ep[1]}; row-domain per chunk (ld done, first half packed, box free, stored, -): {ep[2:7].tolist()} {ep[7:12].tolist()}; coalesced-domain prefetch issue: {pf.tolist()}")
