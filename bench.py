#!/usr/bin/env python
"""Headline benchmark: Conformer-M training step (SpecAugment -> ConvSubSampling -> bridge -> 16 Conformer layers ->
CTC head, backward, gradient all-reduce, optimizer step) on synthetic 80-bin log-mel, 64 x 15 s per GPU,
data-parallel over N GPUs of one node.  Metric (BASELINE.json): audio-seconds per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|infer] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-resident throughput, e2e = the same
step fed from pinned host memory with the loss read back, roofline = dominant kernel (tcgen05 GEMM) measured with
CUDA events in an instrumented step, cpu_baseline = the reference's CPU path on this box's host cores.
`--impl reference` times the reference's own CPU implementation of the path (staged under oracle/_ref by
oracle/stage_ref.py; torchaudio from site-packages) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {   # Conformer(d, heads, ffn, layers, kernel) of SURVEY.md section 8
    "S": dict(d=144, h=4, f=576, L=16, k=31),
    "M": dict(d=256, h=4, f=1024, L=16, k=31),
    "L": dict(d=512, h=8, f=2048, L=17, k=31),
}
NTOK, TGT_LEN = 1024, 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


class HP:
    """The ConvSubSampling / SpecAugment fields of /root/reference/lib/hparams.py at their defaults."""
    conv_sub_1_nodes, conv_sub_1_kernel, conv_sub_1_stride = 512, 7, (2, 2)
    conv_sub_2_nodes, conv_sub_2_kernel, conv_sub_2_stride = 128, 3, (2, 2)
    warping_param_W, warping_ntimes = 1, 1
    frequency_mask_param_F, frequency_mask_ntimes = 5, 2
    time_multiplicity, pm, ps, time_mask_param_T = 2, 0.05, 0.05, 5
    adaptive_multiplicity = adaptive_size = False
    mask_value = 0

    def __init__(self, rows, cols):
        self.input_rows, self.input_cols = rows, cols


def model_flops_fwd(cfg, B, T, Tp, Fo=18, C2=128):
    """Algorithmic forward FLOPs actually EXECUTED per batch (composed conv counted as the 11x11 conv it runs)."""
    d, f, L, k = cfg["d"], cfg["f"], cfg["L"], cfg["k"]
    conv = 2 * 121 * C2 * Fo * Tp
    bridge = 2 * C2 * Fo * d * Tp
    enc = L * Tp * (8 * d * f + 14 * d * d + 4 * Tp * d + 2 * k * d)
    return B * (conv + bridge + enc)


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [c.strip() for c in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ======================================================================================================
# reference arm / CPU baseline: the reference's own modules on the host cores
# ======================================================================================================
def build_reference(cfg, T):
    import torch
    from torch import nn
    kind = "port"
    try:
        from oracle import stage_ref
        if stage_ref.stage():
            sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
            from lib.convsubsampling import ConvSubSampling as RefCS
            from lib.standard.asrnn import ASRNN
            kind = "reference"
    except Exception:
        kind = "port"
    import torchaudio
    hp = HP(80, T)
    hp.device = torch.device("cpu")

    class RefEncoder(nn.Module):
        def __init__(self):
            super().__init__()
            if kind == "reference":
                self.conv_sub_sampling = RefCS(hp, 1, 128)                         # lib/convsubsampling.py (staged copy)
            else:                                                                  # restated: the same two nn.Conv2d
                self.conv_sub_sampling = nn.Sequential(nn.Conv2d(1, 512, 7, stride=2), nn.Conv2d(512, 128, 3, stride=2))
            self.bridge = nn.Linear(128 * 18, cfg["d"])
            self.conformers = torchaudio.models.Conformer(cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], 0.5)
            self.head = nn.Linear(cfg["d"], NTOK)
            self.hp = hp

        def forward(self, x, lens, specaug):
            if specaug:
                if kind == "reference":
                    x = ASRNN.SpecAugment(self, x, lens)                           # asrnn.py:170-192, CPU python loops
                else:
                    from oracle import specaugment_oracle as so
                    x = torch.from_numpy(so.spec_augment(x.numpy(), lens.tolist(), random, so.hp_defaults())[0])
            z = self.conv_sub_sampling(x)                                          # (B,128,18,T')
            z = z.permute(0, 3, 1, 2).flatten(2)
            h = self.bridge(z)
            l2 = ((lens - 7) // 2 + 1 - 3) // 2 + 1
            y, _ = self.conformers(h, l2)
            return self.head(y), l2

    for n in ("time_warping", "frequency_masking", "time_masking"):
        if kind == "reference":
            setattr(RefEncoder, n, getattr(ASRNN, n))
    return RefEncoder(), kind


def reference_step_fn(cfg, B, T, mode):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model, kind = build_reference(cfg, T)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(B, 1, 80, T, generator=g)
    lens = torch.full((B,), T, dtype=torch.int64)
    tgt = torch.randint(1, NTOK, (B, TGT_LEN), generator=g)
    if mode == "train":
        from transformers.optimization import Adafactor
        opt = Adafactor(model.parameters(), lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False)  # runner.py:36
        ctc = torch.nn.CTCLoss(blank=0, zero_infinity=True)                                                   # runner.py:35
        model.train()

        def step():
            opt.zero_grad()
            logits, l2 = model(x, lens, True)
            loss = ctc(logits.log_softmax(-1).transpose(0, 1), tgt, l2, torch.full((B,), TGT_LEN))
            loss.backward()
            opt.step()
            return float(loss)
    else:
        model.eval()

        def step():
            with torch.no_grad():
                return float(model(x, lens, False)[0].sum())
    return step, kind


def run_reference(args, cfg, T, standalone=True):
    B = args.ref_batch
    step, kind = reference_step_fn(cfg, B, T, args.mode)
    warm = args.warmup if standalone else 1
    steps = args.steps if standalone else 2
    for _ in range(warm):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    value = B * T / 100.0 / sec
    sample = f"{B} x {T / 100:.0f} s utterances of the same workload per step ({args.mode}), {steps} timed steps"
    return dict(value=value, unit="audio-s/s", cores=os.cpu_count(), kind=kind, sample=sample), sec


# ======================================================================================================
# our arm
# ======================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--model", default="M", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--ref-batch", type=int, default=2, help="utterances per step of the CPU reference arm")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "hf"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-instrument", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e: copy each batch in front of its step instead of overlapping")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    # stdout carries exactly ONE JSON line: anything native libraries write to fd 1 (NCCL prints its version banner
    # there when NCCL_DEBUG is set) is routed to stderr, and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    cfg = CONFIGS[args.model]
    T = int(round(args.seconds * 100))
    Tp = ((T - 7) // 2 + 1 - 3) // 2 + 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C3 Conformer-{args.model} ({cfg['d']},{cfg['h']},{cfg['f']},{cfg['L']},k{cfg['k']}) "
                + ("training step + SpecAugment" if args.mode == "train" else "inference (eval)")
                + f", {args.batch} x {args.seconds:.0f} s per GPU, 80-bin log-mel @100 fps")

    if args.impl == "reference":
        if rank != 0:
            return
        base, sec = run_reference(args, cfg, T)
        line = dict(metric="audio_seconds_per_second", value=base["value"], unit="audio-s/s", n_gpus=args.gpus,
                    steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=workload, ref_sample=base["sample"]), cpu_baseline=base,
                    e2e=dict(value=base["value"], unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from torch import nn

    import nn_conformer_for_speech_recognition_b200 as b2c
    from nn_conformer_for_speech_recognition_b200 import _lib, dp
    from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()

    B = args.batch
    torch.manual_seed(0)
    random.seed(1234 + rank)
    hp = HP(80, T)
    enc = ConformerEncoder(hp, cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], dropout=0.5 if args.mode == "train" else 0.0,
                           specaug_mode="paper").to(dev)
    import nn_conformer_for_speech_recognition_b200 as b2c_pkg
    head = b2c_pkg.Linear(cfg["d"], NTOK).to(dev)      # CTC head projection on the tcgen05 GEMM (same parameters as nn.Linear)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(B, 1, 80, T, generator=g).pin_memory()
    lens_host = [T] * B
    lens = torch.tensor(lens_host, device=dev)
    tgt = torch.randint(1, NTOK, (B, TGT_LEN), generator=g).to(dev)
    tgt_lens = torch.full((B,), TGT_LEN, device=dev)
    x_dev = x_host.to(dev)
    params = list(enc.parameters()) + list(head.parameters())
    sync = dp.GradSync(enc, head, world=world) if world > 1 else None
    if sync is not None:
        sync.broadcast_parameters()

    from nn_conformer_for_speech_recognition_b200.specaugment import SpecAugPlan
    use_graph = not args.no_graph          # per-rank graphs: the gradient all-reduce runs between backward and optimizer, outside them
    plan = SpecAugPlan(hp, B, dev, mode="paper")
    tgt_lens_host = [TGT_LEN] * B
    out_lens_host = [Tp] * B
    if args.mode == "train":
        enc.train(); head.train()
        if args.optimizer == "fused":
            from nn_conformer_for_speech_recognition_b200.optim import Adafactor
            opt = Adafactor(params, lr=2e-5, beta1=0.9)
        else:
            from transformers.optimization import Adafactor
            opt = Adafactor(params, lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False)
            use_graph = False

        class EncStep(nn.Module):
            """encoder forward as a Module so torch.cuda.make_graphed_callables can capture its forward AND backward"""
            def __init__(self):
                super().__init__()
                self.enc = enc

            def forward(self, xin):
                return self.enc(xin, lens, SpecAugment=True, specaug_plan=plan)[0]

        enc_step = EncStep()

        def device_step(xin):
            """One training step.  The encoder (every kernel of this repo's hot path) is graph-replayed; the CTC head is
            torch (not hot path) and stays eager because torch's CTC copies its length lists host->device internally."""
            opt.zero_grad(set_to_none=True)
            y = enc_step(xin)
            logp = head(y).log_softmax(-1).transpose(0, 1)
            loss = torch.nn.functional.ctc_loss(logp, tgt, out_lens_host, tgt_lens_host, blank=0, zero_infinity=True)
            loss.backward()
            if sync is not None:
                sync.finish()
            opt.step()
            return loss.detach()
    else:
        enc.eval()

        def device_step(xin):
            with torch.no_grad():
                y, _ = enc(xin, lens)
            return y[:, 0, 0].sum()

    x_static = x_dev.clone()
    graph = None
    loss_static = None
    launches_per_step = 0
    graphed = False

    def host_prepare():
        if args.mode == "train":
            plan.draw(80, lens_host)              # SpecAugment draws (CPython `random`, reference order) -> pinned buffer

    if use_graph and args.mode == "train":
        host_prepare()
        l0 = _lib.LAUNCHES
        enc_step = torch.cuda.make_graphed_callables(enc_step, (x_static,), num_warmup_iters=3)
        launches_enc = (_lib.LAUNCHES - l0) // 4          # 3 warm-up passes + 1 capture, each forward + backward
        graphed = True
        launches_per_step = launches_enc + 5              # + the fused Adafactor launches
    elif use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                    # warm-up outside capture: lazy initialisation, caches, allocator pools
                device_step(x_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(graph):
            loss_static = device_step(x_static)
        launches_per_step = _lib.LAUNCHES - l0
        graphed = True

    def step(xin):
        host_prepare()
        if graph is not None:
            if xin is not x_static:
                x_static.copy_(xin, non_blocking=True)
            graph.replay()
            return loss_static
        return device_step(xin)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def timed(fn, n):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b_ in ev:
            flush.zero_()
            a.record()
            fn()
            b_.record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b_) for a, b_ in ev]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_static)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = _lib.LAUNCHES
    ms = timed(lambda: step(x_static), args.steps)
    launches = launches_per_step if graphed else (_lib.LAUNCHES - l0) // args.steps
    barrier()

    # e2e: every step copies one batch from pinned host memory and reads its result back.  The copy of the NEXT batch
    # runs on a copy stream while this step computes (HostPrefetcher, two device buffers) -- what a user's input
    # pipeline does with this package; --no-prefetch serialises copy and step like the reference's loop does.
    from nn_conformer_for_speech_recognition_b200.prefetch import HostPrefetcher
    pre = None if args.no_prefetch else HostPrefetcher(x_host.shape, dev)
    if pre is not None:
        pre.submit(x_host)

    def e2e_step():
        if pre is None:
            x_static.copy_(x_host, non_blocking=True)       # H2D of this step's batch from pinned memory
            return float(step(x_static))                    # D2H read of the step's result (loss / checksum)
        xin = pre.get()                                     # this step's batch (copied while the previous step ran)
        pre.submit(x_host)                                  # H2D of the next step's batch, overlapping this step
        return float(step(xin))
    e2e_step()
    e2e_ms = timed(e2e_step, args.steps)
    barrier()
    clk = clocks.stop()

    tot = torch.tensor([sum(ms), sum(e2e_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_ms, tot_e2e = tot.tolist()
    audio_s = B * args.seconds * world * args.steps
    value = audio_s / (tot_ms / 1e3)
    e2e_val = audio_s / (tot_e2e / 1e3)

    line = dict(metric="audio_seconds_per_second", value=value, unit="audio-s/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=tot_ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=workload, global_batch=B * world, frames_in=T, frames_enc=Tp, parallelism=f"dp{world}",
                            l2="flushed between timed steps (256 MB write)", optimizer=f"Adafactor[{args.optimizer}]",
                            head="Linear(d,1024) on the tcgen05 GEMM; log_softmax + CTC via torch (not hot path, SURVEY 8f#2)",
                            dropout=0.5 if args.mode == "train" else 0.0, specaugment="paper-mode masks + warp",
                            launch=("CUDA graph replay (encoder fwd+bwd captured; CTC head + optimizer eager)" if args.mode == "train" else
                                    "CUDA graph replay") if graphed else "eager launches"),
                clocks=clk,
                e2e=dict(value=e2e_val, unit="audio-s/s", h2d_bytes_per_step=x_host.numel() * 4 * world,
                         d2h_bytes_per_step=4 * world, ms_per_step=tot_e2e / args.steps,
                         input=("next batch copied on a copy stream during the step (HostPrefetcher, 2 device buffers)"
                                if pre is not None else "copied in front of each step")),
                gpu_launches=launches)
    fl = model_flops_fwd(cfg, B, T, Tp) * (3 if args.mode == "train" else 1)
    line["model_tflops_executed"] = fl * world / (tot_ms / args.steps / 1e3) / 1e12

    if rank == 0:
        pk = peaks()
        if not args.no_instrument:
            # instrumented (event-per-launch) step: per-kernel device time, measured live on the launching stream
            from nn_conformer_for_speech_recognition_b200 import conformer as _cf
            _cf._LAYER_CALLS = False               # per-launch events need per-launch calls (eval normally makes one C call per layer)
            prof = _lib.profile_begin()
            host_prepare()
            if args.mode == "train":              # eager, instrumented: one CUDA-event pair per launch
                opt.zero_grad(set_to_none=True)
                y_i = enc(x_static, lens, SpecAugment=True, specaug_plan=plan)[0]
                torch.nn.functional.ctc_loss(head(y_i).log_softmax(-1).transpose(0, 1), tgt, out_lens_host, tgt_lens_host,
                                             blank=0, zero_infinity=True).backward()
                opt.step()
            else:
                device_step(x_static)
            rows = _lib.profile_end(prof)
            agg = {}
            for name, dt, flops, nbytes, _ in rows:
                a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
                a[0] += 1; a[1] += dt; a[2] += flops; a[3] += nbytes
            total = sum(a[1] for a in agg.values()) or 1.0
            top = sorted(agg.items(), key=lambda kv: -kv[1][1])
            line["kernel_time_share"] = {k_: round(v[1] / total, 4) for k_, v in top[:8]}
            line["kernel_time_share_note"] = "eager instrumented step (CUDA-event pair per launch); includes launch gaps"
            # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture of one
            # eager step of this same command (profiles/r01_traffic_train.json, tools/traffic_summary.py); null when the
            # capture does not cover the mode
            traffic = {}
            tp = os.path.join(ROOT, "profiles", "r01_traffic_train.json")
            if args.mode == "train" and args.model == "M" and os.path.exists(tp):
                traffic = json.load(open(tp))

            def traffic_of(prefix):
                sel = [v for k_, v in traffic.items() if prefix in k_]
                n_ = sum(v["launches"] for v in sel)
                return sum(v["dram_bytes_per_launch"] * v["launches"] for v in sel) / n_ if n_ else None
            # roofline of the dominant kernel family: every tcgen05 GEMM launch of one step re-issued back to back
            n_g, us_g, fl_g, _ = _lib.replay(rows, "b2c_gemm_bf16")
            ach = fl_g / (n_g * us_g * 1e-6) / 1e12 if n_g else 0.0
            line["roofline"] = dict(kernel="gemm_tcgen05_kernel: all GEMM launches of one step (fwd, dgrad, wgrad), replayed back to back",
                                    bound="tensor", achieved=ach, peak=pk["sustained"], unit="TFLOP/s", frac=ach / pk["sustained"],
                                    traffic=traffic_of("gemm_tcgen05_kernel"),
                                    traffic_note="mean DRAM bytes per GEMM launch, ncu capture of one eager step (profiles/r01_traffic_train.json)",
                                    peak_source=pk["src"] + " bf16_tflops_sustained (kernels timed inside a long train)",
                                    bound_note="measured limiter is shared-memory bandwidth (operand fills + MMA operand reads + epilogue transpose), see DESIGN.md section 4",
                                    launches=n_g, avg_us=us_g, flops_per_launch=fl_g / max(n_g, 1))
            cname = "b2c_convmod_dw_fwd" if args.mode == "train" else "b2c_convmod_eval_fwd"
            n_c, us_c, _, by_c = _lib.replay(rows, cname)
            if n_c:
                gbs = by_c / (n_c * us_c * 1e-6) / 1e9
                line["roofline_hbm"] = dict(kernel=cname + " (GLU+depthwise[+BN+SiLU]); also FP32-FMA bound: 31 FMA per element",
                                            bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                            traffic=traffic_of("convmod_fwd_kernel"), launches=n_c, avg_us=us_c,
                                            bytes_per_launch=by_c / n_c)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = run_reference(args, cfg, T, standalone=False)
            except Exception as e:      # the baseline is reported, never required for our arm
                line["cpu_baseline"] = dict(value=None, unit="audio-s/s", cores=os.cpu_count(), kind="port",
                                            sample=f"failed: {type(e).__name__}: {e}")
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
