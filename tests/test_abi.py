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


def test_grouped_wgrad_dispatch_is_by_common_k_tile(monkeypatch):
    """Host logic of _lib.gemm_wgrad_grouped (no GPU): one grouped launch when every K divides by 192, or every K by 256;
    anything else (mixed widths, more than four problems) goes out as individual mofo_gemm_wgrad calls."""
    import torch
    from mofo_b200 import _lib

    calls = []

    class FakeLib:
        def mofo_gemm_wgrad_grouped(self, n, *a):
            calls.append(("grouped", n))
            return 0

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())
    monkeypatch.setattr(_lib, "_stream", lambda: None)
    monkeypatch.setattr(_lib, "_ptr", lambda t: None if t is None else t.data_ptr())      # host tensors stand in: nothing is launched
    monkeypatch.setattr(_lib, "gemm_wgrad", lambda dY, X, dW, **kw: calls.append(("single", X.shape[1])))

    def problems(ks, n_out=64, M=128):
        out = []
        for k in ks:
            out.append((torch.zeros(M, n_out, dtype=torch.bfloat16), torch.zeros(M, k, dtype=torch.bfloat16),
                        torch.zeros(n_out, k), None, (0, 0)))
        return out

    _lib.gemm_wgrad_grouped(problems([768, 3072, 768, 768]), 128)          # ViT-B block: 192-wide group
    _lib.gemm_wgrad_grouped(problems([4096, 1024, 1024, 1024]), 128)       # ViT-L block: 256-wide group
    _lib.gemm_wgrad_grouped(problems([512, 2048]), 128)                    # ViT-L decoder
    assert calls == [("grouped", 4), ("grouped", 4), ("grouped", 2)]
    calls.clear()
    _lib.gemm_wgrad_grouped(problems([384, 1024]), 128)                    # no common tile width: 384 % 256 != 0, 1024 % 192 != 0
    _lib.gemm_wgrad_grouped(problems([128]), 128)
    _lib.gemm_wgrad_grouped(problems([768] * 5), 128)                      # more than four problems
    assert calls == [("single", 384), ("single", 1024), ("single", 128)] + [("single", 768)] * 5
    _lib.gemm_wgrad_grouped([], 128)
