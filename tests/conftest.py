import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def emu_lib():
    from tests.emul.build_emul import build
    from planet_heightmap_generation_b200._lib import Library
    return Library(build())


@pytest.fixture(scope="session")
def cuda_lib():
    from planet_heightmap_generation_b200 import build as b
    from planet_heightmap_generation_b200._lib import Library
    if not os.path.exists(b.SO):
        b.build()
    return Library(b.SO)


# Every parity test runs twice: against the host emulation of the kernels (CPU suite) and against
# the real CUDA library (-m gpu).
@pytest.fixture(scope="session", params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_lib")
    if not _has_cuda():
        pytest.fail("gpu test selected but no CUDA device is visible")
    return request.getfixturevalue("cuda_lib")


_planet_cache = {}


def make_planet(oracle, n_cells, seed=42, land=0.3):
    """(SphereMesh, r_xyz, neighborDist, synthetic elevation) — cached per size."""
    key = (n_cells, seed, land)
    if key not in _planet_cache:
        from oracle.mesh_hull import build_sphere_from_points
        from planet_heightmap_generation_b200.sphere import synthetic_elevation
        xyz = oracle.fibonacci_sphere(n_cells, 0.75, seed)
        mesh, xyz = build_sphere_from_points(xyz)
        nd = oracle.neighbor_dist(mesh, xyz)
        elev = synthetic_elevation(xyz, seed, land)
        _planet_cache[key] = (mesh, xyz, nd, elev)
    mesh, xyz, nd, elev = _planet_cache[key]
    return mesh, xyz, nd, elev.copy()


@pytest.fixture(scope="session")
def planet_small(oracle):
    return lambda: make_planet(oracle, 3000)


@pytest.fixture(scope="session")
def planet_medium(oracle):
    return lambda: make_planet(oracle, 20000)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, what
    if a.dtype == np.float32:
        bad = a.view(np.uint32) != b.view(np.uint32)
    else:
        bad = a != b
    if bad.any():
        i = np.nonzero(bad)[0]
        raise AssertionError(f"{what}: {i.size} of {a.size} differ; first at {i[0]}: {a[i[0]]!r} vs {b[i[0]]!r}")
