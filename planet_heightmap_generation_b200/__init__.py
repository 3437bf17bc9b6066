"""B200-native engine for the per-Voronoi-cell hot path of World Orogen
(raguilar011095/planet_heightmap_generation).  See DESIGN.md."""
from ._lib import Library, PlanetB200Error, default_library  # noqa: F401
from .engine import DeviceMesh  # noqa: F401
from .mesh import SphereMesh  # noqa: F401
