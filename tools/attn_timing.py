"""Debug: per-phase cycle stamps of one attention-backward CTA (library must be built with EXTRA=-DB2C_ATTN_TIMING)."""
import sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from nn_conformer_for_speech_recognition_b200._lib import call, ptr, stream_ptr
M, H, T, B = 23872, 4, 373, 64
qkv = torch.randn(M, 3 * H * 64, device="cuda").to(torch.bfloat16)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
att, lse = ops.attention_fwd(qkv, lens, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=1)
datt = torch.randn_like(att)
dqkv = torch.empty_like(qkv)
delta = torch.empty((B, H, T), device="cuda")
dq32 = torch.zeros((3, B * T, H * 64), device="cuda")
for _ in range(2):
    call("b2c_attention_bwd", ptr(qkv), qkv.stride(0), H * 64, 2 * H * 64, ptr(lens), ptr(att), ptr(datt), att.stride(0),
         ptr(lse), ptr(delta), ptr(dq32), ptr(dqkv), B, T, H, 0.125, 0.5, 1, stream_ptr())
torch.cuda.synchronize()
st = dq32.view(-1)[:512].view(torch.int64).cpu().tolist()
for nm, o in (("cta z=1", 128), ("cta z=40", 144)):
    t = st[o:o + 6]
    print(nm, "start->alloc", t[1] - t[0], "alloc->kv", t[2] - t[1], "kv->loop end", t[3] - t[2], "final stores", t[4] - t[3], "dealloc", t[5] - t[4], "total", t[5] - t[0])
for who, off in (("tid0", 0), ("tid300", 64)):
    for i in range(3):
        t = st[off // 1 + i * 8: off + i * 8 + 8] if False else st[(off + i * 8):(off + i * 8 + 8)]
        print(who, "iter", i, "deltas:", [t[k + 1] - t[k] for k in range(7)], "total", t[7] - t[0])
