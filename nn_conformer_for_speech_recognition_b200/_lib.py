"""ctypes binding of libb200conformer.so (the C ABI declared in include/b200conformer.h).

There is NO fallback: if the shared library is missing, or the current device is not sm_100, every entry point
raises.  torch is used only for device memory and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200conformer.so")

BF16, F32 = 0, 1
ACT_NONE, ACT_SILU, ACT_SILU_BWD = 0, 1, 2


class B2CError(RuntimeError):
    pass


class Epilogue(Structure):
    _fields_ = [("bias", c_void_p), ("aux", c_void_p), ("resid", c_void_p), ("out", c_void_p), ("out2", c_void_p),
                ("ld_out", c_int64), ("ld_out2", c_int64), ("ld_aux", c_int64), ("alpha", c_float), ("drop_p", c_float),
                ("drop_seed", c_uint64), ("act", c_int), ("out_dtype", c_int), ("accumulate", c_int), ("reserved", c_int),
                ("colsum", c_void_p)]


class FfnWeights(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln_w", "ln_b", "w1", "b1", "w2", "b2")]


class LayerWeights(Structure):                       # struct b2c_layer_weights, include/b200conformer.h
    _fields_ = ([("ffn1", FfnWeights)] +
                [(n, c_void_p) for n in ("att_ln_w", "att_ln_b", "in_w", "in_b", "out_w", "out_b", "conv_ln_w", "conv_ln_b",
                                         "pw1_w", "pw1_b", "dw_w", "dw_b", "bn_w", "bn_b", "bn_mean", "bn_var", "pw2_w",
                                         "pw2_b")] +
                [("ffn2", FfnWeights), ("final_ln_w", c_void_p), ("final_ln_b", c_void_p)])


_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float
# name -> argtypes ; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "b2c_version": [],
    "b2c_device_check": [],
    "b2c_gemm_bf16": [_P, _L, _I, _P, _L, _I, _L, _L, _L, _I, _I, POINTER(Epilogue), _P],
    "b2c_layernorm_fwd": [_P, _P, _P, _P, _I, _P, _P, _L, _I, _F, _P],
    "b2c_layernorm_bwd": [_P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P],
    "b2c_layernorm_bwd_cast": [_P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P, _F, _F, c_uint64, _P, _P],
    "b2c_layernorm_fwd_pair": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "b2c_layernorm_bwd_pair": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _F, _F, c_uint64, _P, _P],
    "b2c_convmod_eval_fwd": [_P, _P, _P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _P],
    "b2c_convmod_dw_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b2c_convmod_bn_finalize": [_P, _L, _F, _F, _P, _P, _P, _P, _I, _P],
    "b2c_convmod_bn_apply": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "b2c_convmod_bn_finalize_apply": [_P, _P, _L, _F, _F, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "b2c_convmod_bn_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "b2c_convmod_dw_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b2c_attention_bwd": [_P, _L, _L, _L, _P, _P, _P, _L, _P, _P, _P, _P, _I, _I, _I, _F, _F, c_uint64, _P],
    "b2c_attention_fwd": [_P, _L, _L, _L, _P, _P, _L, _P, _I, _I, _I, _F, _F, c_uint64, _P],
    "b2c_attention_rel_fwd": [_P, _L, _L, _L, _L, _P, _L, _L, _P, _P, _L, _P, _I, _I, _I, _F, _F, c_uint64, _P],
    "b2c_attention_rel_bwd": [_P, _L, _L, _L, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _F, c_uint64, _P],
    "b2c_conformer_layer_eval": [_P, POINTER(LayerWeights), _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, ctypes.c_size_t, _P],
    "b2c_convsub_compose": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "b2c_convsub_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "b2c_convsub_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b2c_convsub_compose_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "b2c_specaug_apply": [_P, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _F, _I, _I, _I, _P],
    "b2c_adafactor_step": [_P, _I, _P, _I, _P, _P, _F, _F, _F, _F, _F, _P],
    "b2c_dropout_epoch_bump": [_P],
    "b2c_dropout_epoch_snapshot": [_P, _P],
    "b2c_dropout_epoch_set": [c_uint64, _P],
    "b2c_dropout_epoch_use": [_P],
    "b2c_dropout_export_mask": [_I, _F, c_uint64, _L, _L, _P, _P],
    "b2c_cast_f32_bf16": [_P, _P, _L, _P],
    "b2c_add_f32": [_P, _P, _P, _L, _P],
    "b2c_multi_cast_f32_bf16": [_P, _I, _L, _P],
    "b2c_scale_cast": [_P, _P, _L, _L, _F, _F, c_uint64, _P, _P],
    "b2c_colsum_bf16": [_P, _L, _L, _L, _P, _P],
    "b2c_repack_heads": [_P, _P, _L, _I, _I, _I, _P],
    "b2c_logmel": [_P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "b2c_pad_batch": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "b2c_ffn_fused_eval": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _F, _P],
    "b2c_bn1d_stats": [_P, _P, _L, _I, _P],
    "b2c_bn1d_apply": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "b2c_bn1d_bwd_reduce": [_P, _P, _P, _P, _P, _L, _I, _P],
    "b2c_bn1d_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _L, _I, _P],
    "b2c_ctc_loss": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, ctypes.c_size_t, _P],
    "b2c_lstm_cell_fwd": [_P, _P, _P, _P, _L, _P, _P, _L, _I, _P],
    "b2c_lstm_cell_bwd": [_P, _L, _P, _P, _P, _P, _P, _P, _L, _I, _P],
}
HOST_SIGNATURES = {
    "b2c_dropout_keep_scale": ([c_float], c_float),
    "b2c_conformer_layer_eval_workspace": ([_I, _I, _I, _I, _I], ctypes.c_size_t),
    "b2c_ctc_workspace": ([_I, _I, _I], ctypes.c_size_t),
    "b2c_last_error": ([], c_char_p),
    "b2c_mt19937_seed_host": ([POINTER(c_uint32), c_uint64], None),
    "b2c_mt19937_randint_host": ([POINTER(c_uint32), c_int64, c_int64], c_int64),
}

_lib = None
LAUNCHES = 0          # kernels launched through this binding (bench.py reports it as gpu_launches)
_LAUNCHES_PER_CALL = {"b2c_convsub_compose": 1, "b2c_convsub_compose_bwd": 2, "b2c_attention_bwd": 4, "b2c_attention_rel_bwd": 4, "b2c_conformer_layer_eval": 20, "b2c_adafactor_step": 5, "b2c_ctc_loss": 3, "b2c_logmel": 2, "b2c_version": 0, "b2c_device_check": 0, "b2c_dropout_epoch_use": 0}


def load() -> ctypes.CDLL:
    """Loads the library (once).  Raises B2CError when it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2CError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `make -C nn_conformer_for_speech_recognition_b200/csrc`). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    for name, (args, res) in HOST_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def last_error() -> str:
    return load().b2c_last_error().decode()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_profile = None       # list of (name, start_event, end_event, flops, bytes) while an instrumented step runs


def call(name: str, *args, flops: float = 0.0, nbytes: float = 0.0) -> None:
    """Invoke an int-returning entry point on the current torch stream; raise with the library's message on error.
    flops / nbytes = ALGORITHMIC work of the launch (used only by bench.py's instrumented step)."""
    global LAUNCHES
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(load(), name)(*args)
    if rc != 0:
        kind = ValueError if rc < 0 else B2CError
        raise kind(f"{name} failed (rc={rc}): {last_error()}")
    if _profile is not None:
        e1.record()
        _profile.append((name, e0, e1, flops, nbytes, args))
    LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)


def profile_begin():
    """Start recording a CUDA-event pair around every launch (on the launching stream)."""
    global _profile
    _profile = []
    return _profile


def profile_end(prof):
    """-> [(entry point, milliseconds, flops, bytes, args)]"""
    global _profile
    _profile = None
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b), fl, nb, args) for n, a, b, fl, nb, args in prof]


def replay(rows, name: str, reps: int = 3):
    """Re-issue, back to back and without any Python work in between, every recorded launch of entry point `name`
    (same pointers, shapes and epilogues as in the instrumented step) and time the whole train with ONE CUDA-event pair
    on the launching stream.  Returns (launches per repetition, average microseconds per launch, flops, bytes)."""
    calls = [(r[4], r[2], r[3]) for r in rows if r[0] == name]
    if not calls:
        return 0, 0.0, 0.0, 0.0
    fn = getattr(load(), name)
    for args, _, _ in calls:
        fn(*args)
    torch.cuda.synchronize()
    # Replayed from ONE CUDA graph so that the timed region holds no Python / ctypes launch cost (a ~20 us kernel is
    # about as long as one ctypes call: an eager loop can end up measuring the host).  Captured with capture_begin/end
    # directly: torch.cuda.graph() would empty the allocator cache first and unmap buffers these launches point at.
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            g.capture_begin()
            for args, _, _ in calls:                  # every entry point takes the stream as its last argument
                fn(*(list(args[:-1]) + [side.cuda_stream]))
            g.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        graph = g
    except Exception:
        graph = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if graph is not None:
            graph.replay()
        else:
            for args, _, _ in calls:
                fn(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(calls))
    return len(calls), us, sum(c[1] for c in calls), sum(c[2] for c in calls)


_checked = False


def require_device() -> None:
    """Fail loudly unless a CUDA device of compute capability 10.x is current."""
    global _checked
    if _checked:
        return
    if not torch.cuda.is_available():
        raise B2CError("nn_conformer_for_speech_recognition_b200 needs a B200 (sm_100a) GPU; CUDA is not available "
                       "and there is no CPU fallback")
    call("b2c_device_check")
    _checked = True


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


NVTX = os.environ.get("B2C_NVTX", "0") == "1"


class nvtx:
    """`with nvtx("layer3.mhsa"):` -- an NVTX range around the launches of one module stage when B2C_NVTX=1 (visible in
    Nsight Systems / ncu --nvtx filters); a no-op otherwise (tracing hooks the reference lacks)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


class epoch_scope:
    """`with epoch_scope(word):` -- kernels launched inside read the dropout epoch from `word` (a 1-element int64 device
    tensor written by b2c_dropout_epoch_snapshot) instead of the device counter.  None = no-op."""

    current = None                       # innermost installed word (scopes nest; single-threaded launch path)

    def __init__(self, word):
        self.word = word

    def __enter__(self):
        if self.word is not None:
            self.prev = epoch_scope.current
            epoch_scope.current = self.word
            load().b2c_dropout_epoch_use(self.word.data_ptr())
        return self

    def __exit__(self, *exc):
        if self.word is not None:
            epoch_scope.current = self.prev
            load().b2c_dropout_epoch_use(None if self.prev is None else self.prev.data_ptr())
        return False


def epoch_snapshot(device):
    """New per-call epoch word: bumps the device's step counter and copies it (one capturable launch)."""
    word = torch.empty(1, device=device, dtype=torch.int64)
    call("b2c_dropout_epoch_snapshot", word.data_ptr(), stream_ptr())
    return word
