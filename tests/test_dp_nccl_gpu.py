"""Data-parallel correctness on real GPUs over NCCL ("2/4/8-GPU DP grads == 1-GPU grads on the
concatenated batch").  Needs >= 2 GPUs: the driver's single-GPU `-m gpu` run skips it; run it with
`gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu`.

Each rank takes a contiguous half of the batch, runs the benchmarked training forward/backward (ConformerEncoder + Linear head
+ fused CTC) with `dp.GradSync` attached -- the Conformer stack's gradient arena is all-reduced in layer-group spans from inside
the backward pass, the remaining gradients in `finish()` -- and rank 0 compares EVERY averaged gradient with a single-GPU
run over the whole batch.  BatchNorm runs on its running statistics (eval mode under autograd), because train-mode batch
statistics are per-rank by design (reference semantics: no SyncBN) and would make the two computations differ.
Tolerance: cosine >= 0.9999 and relative L2 <= 2e-2 per parameter (same kernels; only tile / split-K shapes differ with M)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(d=256, h=4, f=1024, L=4, k=31)
B, T = 8, 700


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev):
    import nn_conformer_for_speech_recognition_b200 as b2c
    from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder
    from oracle import ref_pipeline as rp
    torch.manual_seed(0)
    enc = ConformerEncoder(rp.make_hparams(80, T), CFG["d"], CFG["h"], CFG["f"], CFG["L"], CFG["k"], dropout=0.0).to(dev)
    head = b2c.Linear(CFG["d"], 64).to(dev)
    for n, v in enc.state_dict().items():
        if n.endswith("running_mean"):
            v.copy_(0.05 * torch.randn(v.shape, generator=torch.Generator().manual_seed(1)).to(dev))
        if n.endswith("running_var"):
            v.copy_(0.75 + 0.5 * torch.rand(v.shape, generator=torch.Generator().manual_seed(2)).to(dev))
    return enc.eval(), head.eval()


def _data():
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(B, 1, 80, T, generator=g)
    lens = torch.tensor([700, 700, 650, 600, 700, 512, 480, 333])
    tgt = torch.randint(1, 64, (B, 12), generator=g)
    return x, lens, tgt, torch.full((B,), 12)


def _step(enc, head, x, lens, tgt, tl, dev):
    from nn_conformer_for_speech_recognition_b200.ctc import ctc_loss_from_logits
    enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
    y, ol = enc(x.to(dev), lens.to(dev))
    loss = ctc_loss_from_logits(head(y), tgt.to(dev), ol, tl.to(dev))
    loss.backward()
    grads = {"enc." + n: p.grad for n, p in enc.named_parameters()}
    grads.update({"head." + n: p.grad for n, p in head.named_parameters()})
    return loss, grads


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from nn_conformer_for_speech_recognition_b200 import dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    enc, head = _build(dev)
    sync = dp.GradSync(enc, head, world=world).attach(enc.conformers, groups=2)
    sync.broadcast_parameters()
    x, lens, tgt, tl = _data()
    lo, hi = dp.shard_range(B, rank, world)
    loss, grads = _step(enc, head, x[lo:hi], lens[lo:hi], tgt[lo:hi], tl[lo:hi], dev)
    sync.finish()
    torch.cuda.synchronize()
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    if rank == 0:
        enc1, head1 = _build(dev)
        loss1, g1 = _step(enc1, head1, x, lens, tgt, tl, dev)
        rows = []
        for n, g in g1.items():
            a, b = grads[n].double().flatten().cpu().numpy(), g.double().flatten().cpu().numpy()
            nb = np.linalg.norm(b)
            if nb < 1e-12:
                continue
            rows.append((float(np.linalg.norm(a - b) / nb), float(a @ b / (np.linalg.norm(a) * nb + 1e-300)), n))
        rows.sort(reverse=True)
        torch.save(dict(loss_dp=float(torch.stack(losses).mean()), loss_1=float(loss1), rows=rows), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("world", [2])
def test_dp_gradients_equal_single_gpu_gradients(world, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = torch.load(out)
    print(f"loss DP {r['loss_dp']:.6f} vs single {r['loss_1']:.6f}; worst gradients (rel-L2, cosine): {r['rows'][:4]}")
    assert abs(r["loss_dp"] - r["loss_1"]) <= 1e-4 * abs(r["loss_1"])
    for rel, cos, n in r["rows"]:
        assert cos >= 0.9999 and rel <= 2e-2, f"{n}: cosine {cos:.6f} rel-L2 {rel:.3e}"
