"""Reproducibility probe: eager vs eager, graph vs graph, eager vs graph for the benchmarked training step (B=4)."""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_training_gpu as t
from oracle import ref_pipeline as rp
from nn_conformer_for_speech_recognition_b200.specaugment import SpecAugPlan

L = int(os.environ.get("PROBE_L", "16"))
cfg = dict(t.CFG_M, L=L)
B, T = 4, 1500
lens = [1500, 1400, 1300, 900]
_, enc, head = t._pair(cfg, T)
enc.train(); head.train(); enc._b2c_specaug_mode = "paper"
x, lens_t, tgt, tgt_lens = t._batch(B, T, lens)
xd, lens_d = x.cuda(), lens_t.cuda()
plan = SpecAugPlan(enc.hp, B, "cuda", mode="paper")

class Step(torch.nn.Module):
    def __init__(self):
        super().__init__(); self.enc = enc
    def forward(self, xin):
        return self.enc(xin, lens_d, SpecAugment=True, specaug_plan=plan)[0]

def run(fn, seed):
    random.seed(seed); plan.draw(80, lens)
    enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
    y = fn(xd)
    loss = rp.ctc_loss(head(y), tgt.cuda(), enc.output_lengths(lens_d), tgt_lens.cuda())
    loss.backward()
    return loss.item(), y.detach().clone(), {n: p.grad.clone() for n, p in enc.named_parameters()}

def cmp(a, b, tag):
    dy = (a[1] - b[1]).abs()
    rows = []
    for n in a[2]:
        err, sc = (a[2][n] - b[2][n]).abs().max().item(), a[2][n].abs().max().item()
        rows.append((err / (sc + 1e-30), n))
    rows.sort(reverse=True)
    print(f"{tag}: loss {a[0]:.6f} vs {b[0]:.6f}; y max diff {dy.max().item():.3e} mean {dy.mean().item():.3e}; worst grads: " +
          ", ".join(f"{r[0]:.2e} {r[1].replace('conformers.conformer_layers.', 'L')}" for r in rows[:4]))

eager = Step()
e1, e1b = run(eager, 1), run(eager, 1)
cmp(e1, e1b, "eager vs eager")
random.seed(0); plan.draw(80, lens)
graphed = torch.cuda.make_graphed_callables(Step(), (xd,), num_warmup_iters=3)
g1, g1b = run(graphed, 1), run(graphed, 1)
cmp(g1, g1b, "graph vs graph")
cmp(e1, g1, "eager vs graph")
e1c = run(eager, 1)
cmp(e1, e1c, "eager vs eager (after graph)")
