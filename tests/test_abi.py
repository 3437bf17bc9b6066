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


def test_node_addon_type_checks_against_the_header():
    """bindings/node/planet_b200_addon.cc (the N-API shim a maintainer builds with node-gyp) must stay in sync with
    include/planet_b200.h; no Node toolchain exists here, so it is type-checked against a stub <node_api.h>."""
    import subprocess
    src = os.path.join(ROOT, "bindings", "node", "planet_b200_addon.cc")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "bindings", "node", "stub"),
                           "-I" + os.path.join(ROOT, "include"), src])
    text = open(src).read()
    for fn in ("pb_warp_terrain", "pb_smooth_elevation", "pb_erode_composite", "pb_sharpen_ridges", "pb_apply_soil_creep",
               "pb_run_post_processing", "pb_assign_elevation", "pb_compute_wind", "pb_compute_ocean_currents",
               "pb_compute_precipitation", "pb_compute_temperature", "pb_classify_koppen", "pb_climate_get"):
        assert fn in text, fn
