"""Config C5: NST pseudo-labelling sweep -- Conformer-M teacher, eval forward over a pool of
synthetic utterances with lengths U{500..1500} frames (seed 0), length-bucketed batches of 64, contiguous shards per
rank, NO collective on the data path.  Prints one JSON line (audio-seconds of TRUE speech per second, whole job).

    python tools/nst_sweep_bench.py [--utts-per-gpu 1250]          # 1 GPU
    torchrun --nproc-per-node 8 tools/nst_sweep_bench.py            # 8 GPUs: 10 000 utterances in total
"""
import argparse
import json
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_hp  # noqa: E402
from nn_conformer_for_speech_recognition_b200 import _lib, nst  # noqa: E402
from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts-per-gpu", type=int, default=1250)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="M")
    ap.add_argument("--no-bucket", action="store_true", help="reference order (runner.py:270-275): pad to the batch maximum")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    n = args.utts_per_gpu * world
    rng = random.Random(0)
    lengths = [rng.randint(500, 1500) for _ in range(n)]
    cfg = CONFIGS[args.model]
    torch.manual_seed(0)
    enc = ConformerEncoder(make_hp(80, 1500), cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], 0.0).cuda().eval()
    pool = torch.rand(80, 1500, device="cuda")                      # every utterance is a prefix of one resident buffer
    get = lambda i: pool[:, :lengths[i]]                             # noqa: E731
    keep = lambda y, m: int(m)                                       # noqa: E731  labels would be decoded here; keep the length only

    def run():
        return nst.run_sweep(lambda x, l: enc(x, l), get, lengths, args.batch, rank, world, not args.no_bucket, reduce=keep)

    lo, hi = nst.shard_range(n, rank, world)
    plan = nst.plan_sweep(lengths, args.batch, rank, world, not args.no_bucket)
    for sb in plan[:: max(1, len(plan) // 3)]:                      # warm-up a few shapes (TMA descriptor / workspace caches)
        x = torch.zeros(len(sb.indices), 1, 80, sb.max_len, device="cuda")
        with torch.no_grad():
            enc(x, torch.tensor([lengths[i] for i in sb.indices], device="cuda"))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = run()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert sorted(out) == list(range(lo, hi))
    if rank == 0:
        true_s = sum(lengths) / 100.0
        padded = sum(sb.max_len * len(sb.indices) for sb in plan) / max(1, sum(lengths[lo:hi]))
        print(json.dumps(dict(metric="audio_seconds_per_second", value=true_s / (ms.item() / 1e3), unit="audio-s/s",
                              n_gpus=world, utterances=n, ms_total=ms.item(), scaling="weak", dtype="bf16",
                              data="synthetic", collectives_on_data_path=0, gpu_launches=_lib.LAUNCHES - l0,
                              config=dict(workload=f"C5 NST pseudo-labelling sweep, Conformer-{args.model} teacher eval, "
                                          f"{args.utts_per_gpu} utterances per GPU, lengths U{{500..1500}} frames, "
                                          f"{'length-bucketed' if not args.no_bucket else 'in-order'} batches of {args.batch}",
                                          padding_overhead=round(padded, 4)))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
