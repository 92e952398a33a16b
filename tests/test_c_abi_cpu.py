"""CPU-only checks of the drop-in boundary: the shared library builds/loads here (no GPU), exports every symbol the
header declares, the ctypes binding covers them all, and the host-side entry points compute correctly."""
import ctypes
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200conformer.h")


@pytest.fixture(scope="module")
def lib():
    from nn_conformer_for_speech_recognition_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200conformer.h but not exported by the library"


def test_binding_covers_the_header():
    from nn_conformer_for_speech_recognition_b200 import _lib
    bound = set(_lib.SIGNATURES) | set(_lib.HOST_SIGNATURES)
    assert set(declared_symbols()) == bound


def test_version_and_error_string(lib):
    assert lib.b2c_version() == 1
    assert isinstance(lib.b2c_last_error(), bytes)


def test_library_has_no_torch_dependency(lib):
    """The ABI is plain C: the .so must not link against libtorch / libc10 (only the CUDA runtime)."""
    import subprocess
    from nn_conformer_for_speech_recognition_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out


@pytest.mark.parametrize("seed", [0, 1, 42, 1234, 2 ** 31 + 5, 2 ** 40 + 17])
def test_mt19937_host_matches_cpython(lib, seed):
    st = (ctypes.c_uint32 * 625)()
    lib.b2c_mt19937_seed_host(st, seed)
    random.seed(seed)
    for n in [1, 2, 3, 6, 7, 100, 1000, 1501, 2 ** 20 + 3, 2 ** 31 - 1]:
        for _ in range(50):
            assert lib.b2c_mt19937_randint_host(st, 0, n - 1) == random.randint(0, n - 1)
    for _ in range(700):                       # crosses a 624-word regeneration
        assert lib.b2c_mt19937_randint_host(st, -1, 1) == random.randint(-1, 1)


def test_mt19937_empty_range_is_an_error(lib):
    st = (ctypes.c_uint32 * 625)()
    lib.b2c_mt19937_seed_host(st, 0)
    assert lib.b2c_mt19937_randint_host(st, 1, 0) == -2 ** 63     # Python: ValueError("empty range")


def test_sources_avoid_banned_batch_memcpy():
    banned = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync"]
    for dirpath, _, files in os.walk(ROOT):
        if ".git" in dirpath or "gpurun_out" in dirpath or "_ref" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", ".sh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for b in banned:
                    assert b not in text, f"{f} mentions {b}"


def test_dropout_keep_scale_host(lib):
    """1 / keep-probability the fused dropout applies (8-bit decision per element): exact for p = 0.5."""
    assert lib.b2c_dropout_keep_scale(0.0) == 1.0
    assert lib.b2c_dropout_keep_scale(0.5) == 2.0
    assert abs(lib.b2c_dropout_keep_scale(0.1) - 256.0 / (256 - (round(0.1 * 65536) >> 8))) < 1e-6
    assert lib.b2c_dropout_keep_scale(0.25) == pytest.approx(4.0 / 3.0, rel=1e-6)


def test_layer_eval_workspace_and_struct_layout(lib):
    """b2c_conformer_layer_eval's workspace formula and the ctypes mirror of struct b2c_layer_weights (32 pointers)."""
    from nn_conformer_for_speech_recognition_b200 import _lib
    B, T, d, f, H = 64, 373, 256, 1024, 4
    M, Hd = B * T, H * 64
    al = lambda v: (v + 255) // 256 * 256                                # noqa: E731
    want = al(M * d * 2) + al(M * f * 2) + al(M * 3 * Hd * 2) + al(M * Hd * 2) + al(M * 2 * d * 2) + al(M * d * 2)
    assert lib.b2c_conformer_layer_eval_workspace(B, T, d, f, H) == want
    assert ctypes.sizeof(_lib.FfnWeights) == 6 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.LayerWeights) == 32 * ctypes.sizeof(ctypes.c_void_p)
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = hdr[hdr.index("typedef struct b2c_layer_weights {"):hdr.index("} b2c_layer_weights;")]
    assert body.count("const float*") + body.count("const void*") == 20 and body.count("b2c_ffn_weights ffn") == 2
