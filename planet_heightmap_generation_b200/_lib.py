"""ctypes loader for the C ABI in include/planet_b200.h.

The product library is the in-tree CUDA build `libplanet_b200.so`.  There is no CPU path: if the
library is missing this module raises, and pb_context_create fails when no CUDA device exists.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(HERE, "libplanet_b200.so")

POINTER_HOST = 0
POINTER_DEVICE = 1


class PlanetB200Error(RuntimeError):
    pass


class PostParams(C.Structure):
    _fields_ = [("smoothing", C.c_double), ("glacialErosion", C.c_double), ("hydraulicErosion", C.c_double),
                ("thermalErosion", C.c_double), ("ridgeSharpening", C.c_double), ("terrainWarp", C.c_double),
                ("hItersOverride", C.c_int32)]


# every symbol include/planet_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SYMBOLS = {
    "pb_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pb_context_destroy": (None, [_vp]),
    "pb_last_error": (C.c_char_p, []),
    "pb_version": (C.c_char_p, []),
    "pb_set_stream": (C.c_int, [_vp, _vp]),
    "pb_set_pointer_mode": (C.c_int, [_vp, C.c_int]),
    "pb_synchronize": (C.c_int, [_vp]),
    "pb_set_option": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "pb_launch_count": (_i64, []),
    "pb_profile_start": (C.c_int, [_vp, C.c_char_p]),
    "pb_profile_stop": (C.c_int, [_vp, C.c_char_p, _i64]),
    "pb_mesh_create": (C.c_int, [_vp, _i32, _vp, _vp, _vp, C.POINTER(_vp)]),
    "pb_mesh_destroy": (None, [_vp]),
    "pb_mesh_num_regions": (_i32, [_vp]),
    "pb_mesh_num_edges": (_i64, [_vp]),
    "pb_compute_neighbor_dist": (C.c_int, [_vp, _vp]),
    "pb_warp_terrain": (C.c_int, [_vp, _vp, _dbl, _dbl, _vp]),
    "pb_smooth_elevation": (C.c_int, [_vp, _vp, _vp, _i32, _dbl]),
    "pb_priority_flood_carve": (C.c_int, [_vp, _vp, _vp, _dbl, _vp, _vp, _vp]),
    "pb_erode_composite": (C.c_int, [_vp, _vp, _vp, _i32, _dbl, _dbl, _dbl, _i32, _dbl, _dbl, _i32, _dbl]),
    "pb_erode_composite_debug": (C.c_int, [_vp, _vp, _vp, _i32, _dbl, _dbl, _dbl, _i32, _dbl, _dbl, _i32, _dbl,
                                           _i32, _vp, _vp, _vp]),
    "pb_sharpen_ridges": (C.c_int, [_vp, _vp, _vp, _i32, _dbl]),
    "pb_apply_soil_creep": (C.c_int, [_vp, _vp, _vp, _i32, _dbl]),
    "pb_run_post_processing": (C.c_int, [_vp, _vp, C.POINTER(PostParams), _dbl, _vp, _vp, _vp]),
    "pb_last_post_timing": (C.c_int, [_vp, _vp]),
    "pb_smooth_field": (C.c_int, [_vp, _vp, _i32]),
    "pb_assign_elevation": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "pb_climate_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "pb_climate_destroy": (None, [_vp]),
    "pb_compute_wind": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _dbl, _dbl]),
    "pb_compute_ocean_currents": (C.c_int, [_vp, _vp]),
    "pb_compute_precipitation": (C.c_int, [_vp, _vp, _dbl, _dbl]),
    "pb_compute_temperature": (C.c_int, [_vp, _vp, _dbl]),
    "pb_classify_koppen": (C.c_int, [_vp, _vp, _vp]),
    "pb_compute_climate": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _dbl, _dbl, _dbl, _dbl, _vp]),
    "pb_climate_field_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(_i32), C.POINTER(_i64)]),
    "pb_climate_get": (C.c_int, [_vp, C.c_char_p, _vp]),
    "pb_project_coarse_plates": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, C.c_double, _i32, _vp]),
    "pb_smooth_and_reconnect_plates": (C.c_int, [_vp, _vp, _vp, _i32, _i32]),
    "pb_build_super_plates": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pb_generate_coarse_plates": (C.c_int, [_vp, C.c_double, _i32, _i32, C.c_double, C.c_double, _i32, C.POINTER(_vp), _vp, _vp, _vp]),
    "pb_sample_heightmap": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "pb_derive_synthetic_plates": (C.c_int, [_vp, _vp, _vp]),
    "pb_classify_imported_regions": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "pb_region_colors": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "pb_export_map": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pb_mesh_num_triangles": (_i32, [_vp]),
    "pb_mesh_get_triangles": (C.c_int, [_vp, _vp, _vp]),
    "pb_mesh_get_adj_triangles": (C.c_int, [_vp, _vp]),
    "pb_generate_triangle_centers": (C.c_int, [_vp, _vp]),
    "pb_compute_triangle_elevations": (C.c_int, [_vp, _vp, _vp]),
    "pb_generate_fibonacci_sphere": (C.c_int, [_vp, _i32, C.c_double, C.c_double, _vp]),
    "pb_triangulate_sphere": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "pb_mesh_create_from_points": (C.c_int, [_vp, _i32, _vp, C.POINTER(_vp)]),
    "pb_mesh_create_delaunator": (C.c_int, [_vp, _i32, _vp, C.POINTER(_vp)]),
    "pb_mesh_get_adjacency": (C.c_int, [_vp, _vp, _vp]),
    "pb_sweep_shards_create": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "pb_sweep_shards_destroy": (None, [_vp]),
    "pb_sweep_shards_export": (C.c_int, [_vp, _vp]),
    "pb_sweep_shards_connect": (C.c_int, [_vp, _i32, _vp]),
    "pb_sweep_shards_set_min_cells": (C.c_int, [_vp, _i64]),
    "pb_sweep_shards_info": (C.c_int, [_vp, _vp]),
}


class PlateTable(C.Structure):
    _fields_ = [("n", C.c_int32), ("ids", C.c_void_p), ("isOcean", C.c_void_p), ("pole", C.c_void_p), ("omega", C.c_void_p),
                ("density", C.c_void_p)]


class ElevationResult(C.Structure):
    _fields_ = [("r_elevation", C.c_void_p), ("r_stress", C.c_void_p), ("mountain_r", C.c_void_p), ("coastline_r", C.c_void_p),
                ("ocean_r", C.c_void_p), ("debug", C.c_void_p * 12)]


class Library:
    """One loaded copy of the C ABI."""

    def __init__(self, path: str = DEFAULT_SO):
        if not os.path.exists(path):
            raise PlanetB200Error(
                f"{path} not found: build it with `python -m planet_heightmap_generation_b200.build` "
                "(there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)   # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args

    def check(self, status: int):
        if status != 0:
            raise PlanetB200Error(self.dll.pb_last_error().decode("utf-8", "replace"))

    @property
    def version(self) -> str:
        return self.dll.pb_version().decode()


_default = None


def default_library() -> Library:
    global _default
    if _default is None:
        _default = Library(DEFAULT_SO)
    return _default
