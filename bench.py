#!/usr/bin/env python
"""Headline benchmark (BASELINE.json config C3): Conformer-M training step -- SpecAugment -> ConvSubSampling -> bridge ->
dropout -> 16 Conformer layers -> CTC head, backward, gradient all-reduce, optimizer step -- on synthetic 80-bin log-mel,
64 x 15 s per GPU, data-parallel over N GPUs of one node.  Metric: audio-seconds per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|infer] [--model S|M|L] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement):
    value         device-resident throughput (inputs already in HBM), max over ranks
    e2e           the same step fed from pinned host memory every step with the loss read back
    roofline      the dominant kernel family (tcgen05 GEMM), every launch of one step re-issued back to back and timed with
                  CUDA events on the launching stream; achieved = sum 2MNK / time; peak = MEASURED_PEAKS.json
    cpu_baseline  the reference's own CPU path (oracle/ref_pipeline.py: staged lib/convsubsampling.py + ASRNN.SpecAugment +
                  torchaudio Conformer + HF Adafactor + CTC) on this box's host cores, bounded sample (N = 1 only)
    parity        loss / gradient agreement of OUR path with that reference on a shared 2-utterance sample (same state_dict,
                  dropout 0, same SpecAugment draws) -- measured in this very run (N = 1 only)
    secondary     short runs of the other BASELINE.json configs on the same N GPUs: C2 (Conformer-S inference, 64 x 15 s),
                  C4 (Conformer-L inference, 32 x 30 s per GPU), C5 (NST pseudo-labelling sweep shard, 1250 ragged utterances
                  per GPU, no collective)
`--impl reference` times the reference's CPU implementation alone (rank 0; other ranks exit) on the same `config`.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {   # Conformer(d, heads, ffn, layers, kernel) of SURVEY.md section 8
    "S": dict(d=144, h=4, f=576, L=16, k=31),
    "M": dict(d=256, h=4, f=1024, L=16, k=31),
    "L": dict(d=512, h=8, f=2048, L=17, k=31),
}
NTOK, TGT_LEN = 1024, 20
CONFIG_ID = {("S", "infer"): "C2", ("M", "train"): "C3", ("L", "infer"): "C4", ("M", "infer"): "C5-teacher"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


def make_hp(rows, cols):
    """The reference's own HParams object (lib/hparams.py, unchanged -- the config contract of the drop-ins), taken from
    the staged copy; a restatement of the fields the hot path reads only when that copy is missing."""
    from oracle import ref_pipeline
    return ref_pipeline.make_hparams(rows, cols)


def enc_frames(T):
    return ((T - 7) // 2 + 1 - 3) // 2 + 1


def model_flops_fwd(cfg, B, T, Tp, Fo=18, C2=128):
    """Algorithmic forward FLOPs actually EXECUTED per batch (composed conv counted as the 11x11 conv it runs)."""
    d, f, L, k = cfg["d"], cfg["f"], cfg["L"], cfg["k"]
    conv = 2 * 121 * C2 * Fo * Tp
    bridge = 2 * C2 * Fo * d * Tp
    enc = L * Tp * (8 * d * f + 14 * d * d + 4 * Tp * d + 2 * k * d)
    return B * (conv + bridge + enc)


def workload_name(model, mode, batch, seconds):
    cfg = CONFIGS[model]
    cid = CONFIG_ID.get((model, mode), "custom")
    what = "training step (SpecAugment, fwd, bwd, all-reduce, Adafactor)" if mode == "train" else "inference (eval forward)"
    return (f"{cid} Conformer-{model} ({cfg['d']},{cfg['h']},{cfg['f']},{cfg['L']},k{cfg['k']}) {what}, "
            f"{batch} x {seconds:.0f} s per GPU, 80-bin log-mel @100 fps")


def common_config(args, world):
    """`config` of BOTH arms (ours and --impl reference): the same dict, so the driver compares like with like."""
    T = int(round(args.seconds * 100))
    return dict(workload=workload_name(args.model, args.mode, args.batch, args.seconds), global_batch=args.batch * world,
                frames_in=T, frames_enc=enc_frames(T), parallelism=f"dp{world}",
                l2="flushed between timed steps (256 MB write)",
                cpu_sample=f"the CPU reference arm / cpu_baseline processes {args.ref_batch} of the {args.batch} utterances per "
                           f"step (bounded sample of the same workload, all host cores), normalised to audio-s/s")


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
# reference arm / CPU baseline: the reference's own modules on the host cores (oracle/ref_pipeline.py)
# ======================================================================================================
def reference_step_fn(cfg, B, T, mode, state_dict=None, dropout=0.5):
    import torch
    from oracle import ref_pipeline as rp
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = rp.RefEncoder(cfg, T, dropout=dropout if mode == "train" else 0.0)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(B, 1, 80, T, generator=g)
    lens = torch.full((B,), T, dtype=torch.int64)
    tgt = torch.randint(1, NTOK, (B, TGT_LEN), generator=g)
    if mode == "train":
        from transformers.optimization import Adafactor
        opt = Adafactor(model.parameters(), lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False)  # runner.py:36
        model.train()

        def step():
            opt.zero_grad()
            logits, l2 = model(x, lens, "reference")                          # ASRNN.SpecAugment as shipped (CPU python loops)
            loss = rp.ctc_loss(logits, tgt, l2, torch.full((B,), TGT_LEN))
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        model.eval()

        def step():
            with torch.no_grad():
                return float(model(x, lens)[0].sum())
    return step, model.kind


def run_reference(args, cfg, T, standalone=True):
    B = args.ref_batch
    step, kind = reference_step_fn(cfg, B, T, args.mode)
    warm = args.warmup if standalone else 1
    steps = args.steps if standalone else 3
    for _ in range(warm):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    value = B * T / 100.0 / sec
    sample = (f"{B} x {T / 100:.0f} s utterances of the same workload per step ({args.mode}), fp32, torch threads = "
              f"{os.cpu_count()}, {warm} warm-up + {steps} timed steps")
    return dict(value=value, unit="audio-s/s", cores=os.cpu_count(), kind=kind, sample=sample), sec


# ======================================================================================================
# our arm
# ======================================================================================================
class Rig:
    """One model + synthetic batch on this rank's GPU, with the timed-step plumbing shared by the headline and the
    secondary configurations."""

    def __init__(self, args, model, mode, B, seconds, dev, rank, world, graph=True, optimizer="fused"):
        import torch
        from torch import nn
        import nn_conformer_for_speech_recognition_b200 as b2c
        from nn_conformer_for_speech_recognition_b200 import _lib, dp
        from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder
        from nn_conformer_for_speech_recognition_b200.specaugment import SpecAugPlan
        self.torch, self._lib = torch, _lib
        self.model, self.mode, self.B, self.dev, self.rank, self.world = model, mode, B, dev, rank, world
        cfg = self.cfg = CONFIGS[model]
        T = self.T = int(round(seconds * 100))
        self.Tp = enc_frames(T)
        self.seconds = seconds
        torch.manual_seed(0)
        random.seed(1234 + rank)
        hp = self.hp = make_hp(80, T)
        drop = float(getattr(hp, "dropout", 0.5)) if mode == "train" else 0.0          # lib/hparams.py:43
        self.enc = enc = ConformerEncoder(hp, cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], dropout=drop,
                                          specaug_mode="paper").to(dev)
        self.head = head = b2c.Linear(cfg["d"], NTOK).to(dev)    # CTC head projection on the tcgen05 GEMM (nn.Linear parameters)
        g = torch.Generator().manual_seed(1234 + rank)
        self.x_host = torch.rand(B, 1, 80, T, generator=g).pin_memory()
        self.lens_host = [T] * B
        self.lens = lens = torch.tensor(self.lens_host, device=dev)
        self.tgt = tgt = torch.randint(1, NTOK, (B, TGT_LEN), generator=g).to(dev)
        self.x_static = self.x_host.to(dev)
        self.sync = dp.GradSync(enc, head, world=world) if (world > 1 and mode == "train") else None
        if self.sync is not None:
            # the stack's gradient arena is all-reduced in 4 spans of layers from INSIDE the backward pass (captured into the
            # backward graph as a parallel branch), overlapping the backward kernels of the layers below
            self.sync.attach(enc.conformers, groups=int(os.environ.get("B2C_DP_GROUPS", "4")))
            self.sync.broadcast_parameters()
        self.plan = plan = SpecAugPlan(hp, B, dev, mode="paper")
        tgt_lens_host, out_lens_host = [TGT_LEN] * B, [self.Tp] * B
        self.graph, self.loss_static, self.graphed = None, None, False
        self.launches_per_step = 0
        self.opt = None
        use_graph = graph
        if mode == "train":
            enc.train(); head.train()
            params = list(enc.parameters()) + list(head.parameters())
            if optimizer == "fused":
                from nn_conformer_for_speech_recognition_b200.optim import Adafactor
                self.opt = opt = Adafactor(params, lr=2e-5, beta1=0.9)
            else:
                from transformers.optimization import Adafactor
                self.opt = opt = Adafactor(params, lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False)
                use_graph = False

            from nn_conformer_for_speech_recognition_b200.ctc import ctc_loss_from_logits
            fused_ctc = self.fused_ctc = getattr(args, "ctc", "fused") == "fused"
            tgt_i32 = tgt.to(torch.int32)
            out_lens_i32 = torch.full((B,), self.Tp, device=dev, dtype=torch.int32)
            tgt_lens_i32 = torch.full((B,), TGT_LEN, device=dev, dtype=torch.int32)

            def loss_of(y):
                if fused_ctc:     # csrc/ctc.cu: log_softmax + CTC + d/dlogits (asrnn.py:254-256, runner.py:35,143)
                    return ctc_loss_from_logits(head(y), tgt_i32, out_lens_i32, tgt_lens_i32, blank=0, zero_infinity=True)
                logp = head(y).log_softmax(-1).transpose(0, 1)
                return torch.nn.functional.ctc_loss(logp, tgt, out_lens_host, tgt_lens_host, blank=0, zero_infinity=True)
            self.loss_of = loss_of

            class EncStep(nn.Module):
                """forward as a Module so torch.cuda.make_graphed_callables can capture its forward AND backward: the encoder
                and -- with the fused CTC -- the head projection and the loss (everything but the optimizer)"""
                def __init__(self):
                    super().__init__()
                    self.enc, self.head = enc, head

                def forward(self, xin):
                    y = self.enc(xin, lens, SpecAugment=True, specaug_plan=plan)[0]
                    return loss_of(y) if fused_ctc else y

            enc_step = EncStep()
            if use_graph:
                self.host_prepare()
                l0 = _lib.LAUNCHES
                enc_step = torch.cuda.make_graphed_callables(enc_step, (self.x_static,), num_warmup_iters=3)
                self.launches_per_step = (_lib.LAUNCHES - l0) // 4 + 5      # 3 warm-up passes + 1 capture; + the fused Adafactor launches
                self.graphed = True
            sync = self.sync

            def device_step(xin):
                """One training step.  Encoder, head projection and fused CTC loss (every kernel of this repo, INCLUDING the
                bf16 / composed / permuted weight preparation) are graph-replayed, forward and backward; the gradient
                exchange (N > 1) and the fused Adafactor run eagerly behind them."""
                opt.zero_grad(set_to_none=True)
                out = enc_step(xin)
                loss = out if fused_ctc else loss_of(out)
                loss.backward()
                if sync is not None:
                    sync.finish()
                opt.step()
                return loss.detach()
            self.device_step = device_step
        else:
            enc.eval()

            def device_step(xin):
                with torch.no_grad():
                    y, _ = enc(xin, lens)
                return y[:, 0, 0].sum()
            self.device_step = device_step
            if use_graph:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):                # warm-up outside capture: lazy initialisation, caches, allocator pools
                        device_step(self.x_static)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                l0 = _lib.LAUNCHES
                with torch.cuda.graph(self.graph):
                    self.loss_static = device_step(self.x_static)
                self.launches_per_step = _lib.LAUNCHES - l0
                self.graphed = True
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    def host_prepare(self):
        if self.mode == "train":
            self.plan.draw(80, self.lens_host)    # SpecAugment draws (CPython `random`, reference order) -> pinned slot -> H2D

    def step(self, xin):
        if not getattr(self, "_prepared", False):
            self.host_prepare()
        self._prepared = False
        if self.graph is not None:
            if xin is not self.x_static:
                self.x_static.copy_(xin, non_blocking=True)
            self.graph.replay()
            return self.loss_static
        return self.device_step(xin)

    def timed(self, fn, n):
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b_ in ev:
            self.flush.zero_()
            a.record()
            fn()
            b_.record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b_) for a, b_ in ev]

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def measure(self, steps, warmup, prefetch=True):
        """-> dict(ms, e2e_ms, launches): device-resident and end-to-end milliseconds per step, max over ranks."""
        torch, _lib = self.torch, self._lib
        for _ in range(warmup):
            self.step(self.x_static)
        self.barrier()
        l0 = _lib.LAUNCHES
        ms = self.timed(lambda: self.step(self.x_static), steps)
        launches = self.launches_per_step if self.graphed else (_lib.LAUNCHES - l0) // steps
        self.barrier()
        # e2e: every step copies one batch from pinned host memory and reads its result back.  The copy of the NEXT batch
        # runs on a copy stream while this step computes (HostPrefetcher, two device buffers) -- what a user's input
        # pipeline does with this package; --no-prefetch serialises copy and step like the reference's loop does.
        from nn_conformer_for_speech_recognition_b200.prefetch import HostPrefetcher
        pre = HostPrefetcher(self.x_host.shape, self.dev) if prefetch else None
        if pre is not None:
            pre.submit(self.x_host)

        from nn_conformer_for_speech_recognition_b200.prefetch import ResultFetcher
        fetcher = ResultFetcher(self.dev) if prefetch else None
        state = dict(fut=None, losses=[])

        def e2e_step():
            if pre is None:
                self.x_static.copy_(self.x_host, non_blocking=True)   # H2D of this step's batch from pinned memory
                state["losses"].append(float(self.step(self.x_static)))   # D2H read of the step's result (loss / checksum)
                return
            xin = pre.get()                                           # this step's batch (copied while the previous step ran)
            pre.submit(self.x_host)                                   # H2D of the next step's batch, overlapping this step
            loss = self.step(xin)
            nxt = fetcher.fetch(loss)                                 # D2H of THIS step's loss, enqueued in stream order
            self.host_prepare()                                       # the NEXT step's SpecAugment draws (host, CPython `random`) are made
            self._prepared = True                                     # while the GPU runs this step -- still once per step, inside the timed region
            if state["fut"] is not None:
                state["losses"].append(state["fut"].value())          # host reads the PREVIOUS step's loss while this step runs
            state["fut"] = nxt

        def e2e_drain():
            if state["fut"] is not None:
                state["losses"].append(state["fut"].value())
                state["fut"] = None
        e2e_step()
        e2e_drain()
        # ONE event pair around all K steps (host launch gaps between steps are inside it); the L2-flush fills between steps
        # are timed by their own event pairs and subtracted.  The last step's loss is read before the clock stops.
        torch.cuda.synchronize()
        n0 = len(state["losses"])
        fl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
        for f0, f1 in fl:
            f0.record()
            self.flush.zero_()
            f1.record()
            e2e_step()
        e2e_drain()
        ev_b.record()
        torch.cuda.synchronize()
        assert len(state["losses"]) - n0 == steps, "every step's result must have been read back inside the timed region"
        e2e_ms = [(ev_a.elapsed_time(ev_b) - sum(f0.elapsed_time(f1) for f0, f1 in fl)) / steps] * steps
        self.barrier()
        tot = torch.tensor([sum(ms), sum(e2e_ms)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(tot, op=torch.distributed.ReduceOp.MAX)
        tot_ms, tot_e2e = tot.tolist()
        return dict(ms=tot_ms / steps, e2e_ms=tot_e2e / steps, launches=launches, prefetch=pre is not None)

    def instrumented_rows(self):
        """One eager step with a CUDA-event pair around every launch of this library (on the launching stream)."""
        torch, _lib = self.torch, self._lib
        from nn_conformer_for_speech_recognition_b200 import conformer as _cf
        saved = _cf._LAYER_CALLS
        _cf._LAYER_CALLS = False               # per-launch events need per-launch calls (eval normally makes one C call per layer)
        gs = getattr(self.enc.conformers, "_grad_sync", None)
        self.enc.conformers._grad_sync = None  # this step runs on rank 0 ALONE: no collective may be issued from its backward
        prof = _lib.profile_begin()
        self.host_prepare()
        if self.mode == "train":
            self.opt.zero_grad(set_to_none=True)
            y_i = self.enc(self.x_static, self.lens, SpecAugment=True, specaug_plan=self.plan)[0]
            self.loss_of(y_i).backward()
            self.opt.step()
        else:
            with torch.no_grad():
                self.enc(self.x_static, self.lens)
        rows = _lib.profile_end(prof)
        _cf._LAYER_CALLS = saved
        self.enc.conformers._grad_sync = gs
        return rows

    def gemm_roofline(self, rows, pk):
        n_g, us_g, fl_g, by_g = self._lib.replay(rows, "b2c_gemm_bf16")
        ach = fl_g / (n_g * us_g * 1e-6) / 1e12 if n_g else 0.0
        r = dict(achieved=ach, peak=pk["sustained"], frac=ach / pk["sustained"], launches=n_g, avg_us=us_g,
                 flops_per_launch=fl_g / max(n_g, 1))
        if n_g and by_g:
            # the same launches against the OTHER roof: these GEMMs are skinny (K = 256..1024 with [M, 1024] bf16 or
            # [M, 256] fp32 epilogue operands), so their algorithmic HBM bytes take longer at the measured copy bandwidth
            # than their flops at the measured tensor peak -- the roof that binds is the one with the larger floor
            gbs = by_g / (n_g * us_g * 1e-6) / 1e9
            t_hbm, t_tc = by_g / (pk["hbm"] * 1e9), fl_g / (pk["sustained"] * 1e12)
            r["hbm_side"] = dict(bytes_per_launch=by_g / n_g, achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                 floor_ms_hbm=t_hbm * 1e3, floor_ms_tensor=t_tc * 1e3,
                                 binding_roof="hbm" if t_hbm > t_tc else "tensor",
                                 note="algorithmic bytes: operands, epilogue operands and outputs each cross HBM once")
        return r


def lib_hash():
    """Identity of the kernel code: sha256 over the library's SOURCES (csrc/*, the C header), not over the .so -- nvcc embeds
    per-build identifiers, so two builds of the same sources differ in a few dozen bytes and a binary hash would never match
    on a box that rebuilt the library."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "nn_conformer_for_speech_recognition_b200", "csrc", "*.cu")) +
                   glob.glob(os.path.join(ROOT, "nn_conformer_for_speech_recognition_b200", "csrc", "*.cuh")) +
                   glob.glob(os.path.join(ROOT, "nn_conformer_for_speech_recognition_b200", "csrc", "*.cpp")) +
                   glob.glob(os.path.join(ROOT, "nn_conformer_for_speech_recognition_b200", "csrc", "Makefile")) +
                   glob.glob(os.path.join(ROOT, "include", "*.h")))
    h = hashlib.sha256()
    try:
        for f in files:
            h.update(os.path.basename(f).encode())
            h.update(open(f, "rb").read())
    except OSError:
        return None
    return h.hexdigest()[:16] if files else None


def load_traffic(mode, model):
    """DRAM bytes per launch from the committed ncu capture -- used ONLY when it was taken from this very kernel code (the
    capture records the sha256 of the library's sources, `lib_hash()`; see tools/traffic_capture.sh)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic_train.json")
    if not (mode == "train" and model == "M" and os.path.exists(p)):
        return {}, "no ncu traffic capture for this mode/model"
    d = json.load(open(p))
    if d.get("src_sha256_16") != lib_hash():
        return {}, f"profiles/r02_traffic_train.json was captured from other kernel sources ({d.get('src_sha256_16')}); not used"
    return d.get("kernels", {}), "mean DRAM bytes per launch, ncu capture of one eager step of these kernel sources (profiles/r02_traffic_train.json)"


def parity_check(rig):
    """OUR training step against the reference pipeline on a shared sample, in this run: same state_dict (taken from the
    model just benchmarked), dropout 0 on both sides, paper-mode SpecAugment under the same `random.seed`, 2 x 15 s."""
    import torch
    from oracle import ref_pipeline as rp
    enc, head = rig.enc, rig.head
    n = 2
    x = rig.x_host[:n].clone()
    lens = [rig.T, rig.T - 163]
    g = torch.Generator().manual_seed(99)
    tgt = torch.randint(1, NTOK, (n, TGT_LEN), generator=g)
    tl = torch.full((n,), TGT_LEN)
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    sd.update({"head." + k: v.detach().cpu() for k, v in head.state_dict().items()})
    ref = rp.RefEncoder(rig.cfg, rig.T, dropout=0.0)
    ref.load_state_dict(sd)
    ref.train()
    random.seed(4321)
    logits_r, l2 = ref(x, torch.tensor(lens), "paper")
    loss_r = rp.ctc_loss(logits_r, tgt, l2, tl)
    loss_r.backward()
    saved = (enc.conformers.dropout_p, enc.bridge_dropout_p)
    enc.conformers.dropout_p = enc.bridge_dropout_p = 0.0
    enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
    bn_state = {k: v.clone() for k, v in enc.state_dict().items() if "running_" in k or "num_batches" in k}
    random.seed(4321)
    y, ol = enc(x.to(rig.dev), torch.tensor(lens, device=rig.dev), SpecAugment=True, lens_host=lens)
    loss = rp.ctc_loss(head(y), tgt.to(rig.dev), ol, tl.to(rig.dev))
    loss.backward()
    ours = {k: p.grad.detach().float().cpu() for k, p in enc.named_parameters()}
    ours.update({"head." + k: p.grad.detach().float().cpu() for k, p in head.named_parameters()})
    refg = {k: p.grad for k, p in ref.named_parameters()}
    num = sum(float(((ours[k] - refg[k]) ** 2).sum()) for k in refg) ** 0.5
    den = sum(float((refg[k] ** 2).sum()) for k in refg) ** 0.5
    cos = []
    for k in refg:
        a, b = ours[k].flatten().double(), refg[k].flatten().double()
        if float(b.norm()) > 1e-9 and "conv_module.sequential.2.bias" not in k:
            cos.append((float(a @ b / (a.norm() * b.norm() + 1e-300)), k))
    enc.conformers.dropout_p, enc.bridge_dropout_p = saved
    enc.load_state_dict(bn_state, strict=False)
    enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
    worst = min(cos)
    return dict(sample="2 x 15 s (lengths 1500 / 1337 frames), this run's state_dict in both, dropout 0, paper-mode SpecAugment "
                       "under random.seed(4321); reference = oracle/ref_pipeline.RefEncoder fp32 on the host",
                reference_kind=ref.kind, loss_ours=float(loss), loss_reference=float(loss_r),
                loss_rel_diff=abs(float(loss) - float(loss_r)) / abs(float(loss_r)),
                grad_global_rel_l2=num / den, min_param_grad_cosine=worst[0], min_cosine_param=worst[1],
                n_params_compared=len(cos), tolerance="bf16 operands / fp32 accumulate: loss rel <= 2e-3, per-parameter cosine >= 0.995")


def run_secondary(args, dev, rank, world, pk):
    """Short runs of the other BASELINE.json configurations on the same GPUs (all ranks run their own shard / replica;
    no collective on the data path; time = max over ranks)."""
    import torch
    import torch.distributed as dist
    from nn_conformer_for_speech_recognition_b200 import _lib, nst
    out = {}
    for cid, model, B, seconds in (("C2", "S", 64, 15.0), ("C4", "L", 32, 30.0)):
        rig = Rig(args, model, "infer", B, seconds, dev, rank, world, graph=True)
        clocks = ClockSampler(dev.index)
        clocks.start()
        m = rig.measure(max(5, min(args.steps, 10)), 3)
        clk = clocks.stop()
        audio = B * seconds * world
        entry = dict(workload=workload_name(model, "infer", B, seconds), n_gpus=world, ms=m["ms"],
                     value=audio / (m["ms"] / 1e3), unit="audio-s/s", e2e_ms=m["e2e_ms"], e2e_value=audio / (m["e2e_ms"] / 1e3),
                     gpu_launches=m["launches"], launch="CUDA graph replay", dtype="bf16", clocks=clk,
                     model_tflops_executed=model_flops_fwd(rig.cfg, B, rig.T, rig.Tp) * world / (m["ms"] / 1e3) / 1e12)
        if rank == 0:
            r = rig.gemm_roofline(rig.instrumented_rows(), pk)
            entry["roofline"] = dict(kernel="gemm_tcgen05_kernel (all GEMM launches of one batch)", bound="tensor", unit="TFLOP/s", **r)
        out[cid] = entry
        del rig
        torch.cuda.empty_cache()
    # ---- C5: NST pseudo-labelling sweep (runner.py:253-281), Conformer-M teacher, ragged lengths, eager (each bucket
    # shape occurs once, so graphs do not pay), one C call per layer
    from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder
    per_gpu = args.nst_utts
    n = per_gpu * world
    rng = random.Random(0)
    lengths = [rng.randint(500, 1500) for _ in range(n)]
    cfg = CONFIGS["M"]
    torch.manual_seed(0)
    enc = ConformerEncoder(make_hp(80, 1500), cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], 0.0).to(dev).eval()
    pool = torch.rand(80, 1500, device=dev)
    get = lambda i: pool[:, :lengths[i]]                             # noqa: E731
    keep = lambda y, mlen: int(mlen)                                 # noqa: E731  labels would be decoded here
    plan = nst.plan_sweep(lengths, 64, rank, world, True)
    for sb in plan[:: max(1, len(plan) // 3)]:                       # warm a few shapes (tensor-map / workspace caches)
        xw = torch.zeros(len(sb.indices), 1, 80, sb.max_len, device=dev)
        with torch.no_grad():
            enc(xw, torch.tensor([lengths[i] for i in sb.indices], device=dev))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index)
    clocks.start()
    l0 = _lib.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = nst.run_sweep(lambda x, l: enc(x, l), get, lengths, 64, rank, world, True, device=dev, reduce=keep)
    b.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    lo, hi = nst.shard_range(n, rank, world)
    assert sorted(res) == list(range(lo, hi))
    out["C5"] = dict(workload=f"C5 NST pseudo-labelling sweep, Conformer-M teacher eval, {per_gpu} utterances per GPU "
                              f"({n} in total), lengths U{{500..1500}} frames (seed 0), length-bucketed batches of 64, "
                              "contiguous shard per rank, no collective", n_gpus=world, ms_total=ms.item(),
                     value=sum(lengths) / 100.0 / (ms.item() / 1e3), unit="audio-s/s (true speech seconds)",
                     gpu_launches=_lib.LAUNCHES - l0, launch="eager, one C call per layer", dtype="bf16", clocks=clk,
                     collectives_on_data_path=0)
    return out


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
    ap.add_argument("--ref-batch", type=int, default=8, help="utterances per step of the CPU reference arm")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "hf"])
    ap.add_argument("--ctc", default="fused", choices=["fused", "torch"], help="fused = csrc/ctc.cu (default); torch = ATen log_softmax + ctc_loss")
    ap.add_argument("--nst-utts", type=int, default=1250, help="C5 secondary: utterances per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-instrument", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
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
    Tp = enc_frames(T)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = common_config(args, max(world, args.gpus) if args.impl == "reference" else world)

    if args.impl == "reference":
        if rank != 0:
            return
        base, sec = run_reference(args, cfg, T)
        line = dict(metric="audio_seconds_per_second", value=base["value"], unit="audio-s/s", n_gpus=args.gpus,
                    steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=config, cpu_baseline=base,
                    e2e=dict(value=base["value"], unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        emit(line)
        return

    import torch
    import torch.distributed as dist

    from nn_conformer_for_speech_recognition_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()

    rig = Rig(args, args.model, args.mode, args.batch, args.seconds, dev, rank, world, graph=not args.no_graph,
              optimizer=args.optimizer)
    clocks = ClockSampler(local_rank)
    clocks.start()
    m = rig.measure(args.steps, args.warmup, prefetch=not args.no_prefetch)
    clk = clocks.stop()
    B = args.batch
    audio_s = B * args.seconds * world
    value = audio_s / (m["ms"] / 1e3)
    e2e_val = audio_s / (m["e2e_ms"] / 1e3)
    details = dict(head=("Linear(d,1024) on the tcgen05 GEMM + fused log_softmax/CTC kernels (csrc/ctc.cu)" if getattr(rig, "fused_ctc", False)
                         else "Linear(d,1024) on the tcgen05 GEMM; log_softmax + CTC via torch"),
                   launch=(("CUDA graph replay (encoder + head + CTC fwd+bwd incl. weight preparation captured; optimizer eager)"
                            if args.mode == "train" else "CUDA graph replay") if rig.graphed else "eager launches"))
    if args.mode == "train":
        details.update(optimizer=f"Adafactor[{args.optimizer}] lr 2e-5 beta1 0.9 (runner.py:36)",
                       dropout=float(getattr(rig.hp, "dropout", 0.5)), specaugment="paper-mode masks + warp",
                       hparams=type(rig.hp).__module__ + "." + type(rig.hp).__name__)
    line = dict(metric="audio_seconds_per_second", value=value, unit="audio-s/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=m["ms"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic", config=config, details=details, clocks=clk,
                e2e=dict(value=e2e_val, unit="audio-s/s", h2d_bytes_per_step=rig.x_host.numel() * 4 * world,
                         d2h_bytes_per_step=4 * world, ms_per_step=m["e2e_ms"],
                         input=("next batch copied on a copy stream during the step (HostPrefetcher, 2 device buffers)"
                                if m["prefetch"] else "copied in front of each step"),
                         result=("loss of step n copied to pinned memory in stream order and read by the host while step n+1 "
                                 "runs (ResultFetcher); all K losses are read before the clock stops"
                                 if m["prefetch"] else "loss.item() behind every step"),
                         timing="one CUDA-event pair around all K steps (host launch gaps included) minus the timed L2-flush fills"),
                gpu_launches=m["launches"])
    fl = model_flops_fwd(cfg, B, T, Tp) * (3 if args.mode == "train" else 1)
    line["model_tflops_executed"] = fl * world / (m["ms"] / 1e3) / 1e12

    pk = peaks()
    if rank == 0 and not args.no_instrument:
        rows = rig.instrumented_rows()
        agg = {}
        for name, dt, flops, nbytes, _ in rows:
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += dt; a[2] += flops; a[3] += nbytes
        total = sum(a[1] for a in agg.values()) or 1.0
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])
        line["kernel_time_share"] = {k_: round(v[1] / total, 4) for k_, v in top[:8]}
        line["kernel_time_share_note"] = "eager instrumented step (CUDA-event pair per launch); includes launch gaps"
        traffic, traffic_note = load_traffic(args.mode, args.model)

        def traffic_of(prefix):
            sel = [v for k_, v in traffic.items() if prefix in k_]
            n_ = sum(v["launches"] for v in sel)
            return sum(v["dram_bytes_per_launch"] * v["launches"] for v in sel) / n_ if n_ else None
        r = rig.gemm_roofline(rows, pk)
        kname = "gemm_tcgen05_kernel: all GEMM launches of one step (fwd, dgrad, wgrad), replayed back to back"
        hs = r.pop("hbm_side", None)
        if hs and hs["binding_roof"] == "hbm":
            # the roof that binds this family is HBM: its algorithmic bytes need longer at the measured copy bandwidth than
            # its flops at the measured tensor peak.  The tensor-side figure (round 1's headline fraction) stays beside it.
            line["roofline"] = dict(kernel=kname, bound="hbm", unit="GB/s", achieved=hs["achieved"], peak=hs["peak"],
                                    frac=hs["frac"], traffic=traffic_of("gemm_tcgen05_kernel"), traffic_note=traffic_note,
                                    peak_source=pk["src"] + " hbm_gbs (copy bandwidth)", launches=r["launches"],
                                    avg_us=r["avg_us"], bytes_per_launch=hs["bytes_per_launch"],
                                    floor_ms_hbm=hs["floor_ms_hbm"], floor_ms_tensor=hs["floor_ms_tensor"],
                                    bytes_note=hs["note"],
                                    tensor_side=dict(bound="tensor", unit="TFLOP/s", achieved=r["achieved"], peak=r["peak"],
                                                     frac=r["frac"], flops_per_launch=r["flops_per_launch"],
                                                     peak_source=pk["src"] + " bf16_tflops_sustained"))
        else:
            line["roofline"] = dict(kernel=kname, bound="tensor", unit="TFLOP/s", traffic=traffic_of("gemm_tcgen05_kernel"),
                                    traffic_note=traffic_note, hbm_side=hs,
                                    peak_source=pk["src"] + " bf16_tflops_sustained (kernels timed inside a long train)", **r)
        cname = "b2c_convmod_dw_fwd" if args.mode == "train" else "b2c_convmod_eval_fwd"
        n_c, us_c, _, by_c = _lib.replay(rows, cname)
        if n_c:
            gbs = by_c / (n_c * us_c * 1e-6) / 1e9
            line["roofline_hbm"] = dict(kernel=cname + " (GLU+depthwise[+BN+SiLU]); also FP32-FMA bound: 31 FMA per element",
                                        bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                        traffic=traffic_of("convmod_fwd_kernel"), launches=n_c, avg_us=us_c,
                                        bytes_per_launch=by_c / n_c)
    if rank == 0 and world == 1 and args.mode == "train" and not args.no_parity:
        try:
            line["parity"] = parity_check(rig)
        except Exception as e:
            line["parity"] = dict(error=f"{type(e).__name__}: {e}")
    if not args.no_secondary:
        del rig
        torch.cuda.empty_cache()
        try:
            sec = run_secondary(args, dev, rank, world, pk)
        except Exception as e:          # secondary lines never take the headline down
            sec = dict(error=f"{type(e).__name__}: {e}")
        line["secondary"] = sec
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = run_reference(args, cfg, T, standalone=False)
            except Exception as e:      # the baseline is reported, never required for our arm
                line["cpu_baseline"] = dict(value=None, unit="audio-s/s", cores=os.cpu_count(), kind="port",
                                            sample=f"failed: {type(e).__name__}: {e}")
        emit(line)
    if world > 1:
        # The backward CUDA graphs hold captured NCCL collectives; tearing the communicator down underneath them can block
        # (observed: destroy_process_group() never returned at N = 2).  All results are out: meet once more, then leave.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
