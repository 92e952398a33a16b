"""Fused Adafactor on the Conformer-M parameter set: CUDA-event time of one step (L2 flushed), and under
`ncu --metrics gpu__time_duration.sum -k regex:adafactor` the per-kernel split."""
import sys
import torch
sys.path.insert(0, ".")
import nn_conformer_for_speech_recognition_b200 as b2c
from nn_conformer_for_speech_recognition_b200.optim import Adafactor

from bench import make_hp as HP
enc = b2c.ConformerEncoder(HP(80, 1500), 256, 4, 1024, 16, 31).cuda()
head = torch.nn.Linear(256, 1024).cuda()
params = [p for p in list(enc.parameters()) + list(head.parameters()) if p.requires_grad]
arena = torch.randn(sum((p.numel() + 3) // 4 * 4 for p in params), device="cuda") * 0.01
o = 0
for p in params:
    p.grad = arena[o:o + p.numel()].view(p.shape)
    o += (p.numel() + 3) // 4 * 4
opt = Adafactor(params, lr=2e-5, beta1=0.9)
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts = []
for _ in range(10):
    flush.zero_()
    ev[0].record()
    opt.step()
    ev[1].record()
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
ts.sort()
n = sum(p.numel() for p in params)
print(f"adafactor step: {n/1e6:.1f} M params, median {ts[len(ts)//2]:.0f} us (min {ts[0]:.0f}), "
      f"{28*n/ts[len(ts)//2]/1e3:.0f} GB/s of the 28 B/param model")
