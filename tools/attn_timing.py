"""Debug: per-phase cycle stamps of one attention-backward CTA (library built with EXTRA=-DB2C_ATTN_TIMING, selected
with B2C_DBG_LIB=<file name>; the stamps land in an extra partial-dQ slab behind the real ones)."""
import os, sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops, _lib as _l
_l.LIB_PATH = _l.LIB_PATH.replace("libb200conformer.so", os.environ.get("B2C_DBG_LIB", "libb200conformer.so"))
from nn_conformer_for_speech_recognition_b200._lib import call, ptr, stream_ptr
M, H, T, B = 23872, 4, 373, 64
qkv = torch.randn(M, 3 * H * 64, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
att, lse = ops.attention_fwd(qkv, lens, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=1)
datt = torch.randn_like(att)
dqkv = torch.empty_like(qkv)
delta = torch.empty((B, H, T), device="cuda")
dq32 = torch.zeros((4, B * T, H * 64), device="cuda")
for _ in range(2):
    call("b2c_attention_bwd", ptr(qkv), qkv.stride(0), H * 64, 2 * H * 64, ptr(lens), ptr(att), ptr(datt), att.stride(0),
         ptr(lse), ptr(delta), ptr(dq32), ptr(dqkv), B, T, H, 0.125, 0.5, 1, stream_ptr())
torch.cuda.synchronize()
st = dq32[3].view(-1)[:512].view(torch.int64).cpu().tolist()
t0 = st[0]
print("kernel stamps (start, alloc, -, loop end, stores, dealloc):", [x - t0 if x else None for x in st[:6]])
for i in range(3):
    print("math  tid0 iter", i, "(before S wait, S ready, tile math done, dQ(i-1) drained, arrived):", [x - t0 for x in st[16 + i * 8:16 + i * 8 + 5]])
    print("issuer     iter", i, "(before tile wait, tile ready, S/dP(i+1) issued, 24 MMAs issued, Q(i+2) loaded):", [x - t0 for x in st[64 + i * 8:64 + i * 8 + 5]])
