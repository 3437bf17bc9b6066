"""The C-ABI library loads without a GPU and exports every symbol include/planet_b200.h declares; the host
mirror refuses to run without the CUDA library (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "planet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported():
    from planet_heightmap_generation_b200 import build as b
    from planet_heightmap_generation_b200._lib import SYMBOLS, Library
    names = _declared()
    assert len(names) >= 30
    assert set(names) == set(SYMBOLS), set(names) ^ set(SYMBOLS)
    so = b.build()
    lib = Library(so)                       # resolves every symbol; raises AttributeError otherwise
    dll = ctypes.CDLL(so)
    for n in names:
        assert hasattr(dll, n), n
    assert "cuda" in lib.version


def test_no_cpu_fallback(tmp_path):
    from planet_heightmap_generation_b200._lib import Library, PlanetB200Error
    with pytest.raises(PlanetB200Error):
        Library(str(tmp_path / "missing.so"))


def test_context_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from planet_heightmap_generation_b200 import build as b
    from planet_heightmap_generation_b200._lib import Library
    lib = Library(b.build())
    ctx = ctypes.c_void_p()
    assert lib.dll.pb_context_create(0, ctypes.byref(ctx)) != 0
    assert b"CUDA" in lib.dll.pb_last_error() or b"cuda" in lib.dll.pb_last_error()
