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
