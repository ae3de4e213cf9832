"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol that
include/mofo_b200.h declares (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mofo_b200 import build, _lib
    build.build()
    return _lib


def declared_symbols():
    with open(os.path.join(ROOT, "include", "mofo_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mofo_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    cdll = lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/mofo_b200.h but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in mofo_b200/_lib.py"
    assert sorted(lib.SIGNATURES) == names


def test_version_and_error_string(lib):
    cdll = lib.load()
    assert cdll.mofo_version() == 200          # MOFO_B200_VERSION (include/mofo_b200.h)
    assert isinstance(cdll.mofo_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # invalid arguments are rejected on the host before any CUDA call
    cdll = lib.load()
    rc = cdll.mofo_gemm_tn(None, 0, None, 0, 0, 0, 0, 0, None, None, 0, None, 0, None, None, 0, 0, None, 0, None, 0, None, None)
    assert rc == -1 and b"null" in cdll.mofo_last_error()
    rc = cdll.mofo_tube_mask_bb(None, None, 0, 0, 8, 14, 14, 176, 0.75, None, None, None, None, None)
    assert rc == -1


def test_missing_library_fails_loudly(monkeypatch, lib):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libmofo_sm100.so")
    with pytest.raises(lib.MofoError):
        lib.load()
