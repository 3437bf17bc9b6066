"""Device-resident mesh + context: the object that stands in for the reference's `mesh` argument.

Every hot-path function of the reference takes `mesh` ({numRegions, adjOffset, adjList},
js/sphere-mesh.js:94-146) and `r_xyz`.  `DeviceMesh` uploads both once and keeps them (and
neighborDist) in HBM, like the worker's retained state `W` (js/planet-worker.js:277-292).

Array arguments may be numpy arrays (host buffers: copied in and out by the C ABI, the call blocks)
or torch CUDA tensors (already resident: work is only enqueued on torch's current stream).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Library, PlanetB200Error, PostParams, default_library

_NP = {"f32": np.float32, "i32": np.int32, "u8": np.uint8}


def _is_torch_cuda(a) -> bool:
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


class DeviceMesh:
    def __init__(self, mesh, r_xyz, device: int = 0, lib: Library | None = None):
        self.lib = lib or default_library()
        self.numRegions = int(mesh.numRegions)
        self.adjOffset = np.ascontiguousarray(mesh.adjOffset, np.int32)
        self.adjList = np.ascontiguousarray(mesh.adjList, np.int32)
        self.r_xyz = np.ascontiguousarray(r_xyz, np.float32).reshape(-1)
        if self.adjOffset.shape[0] != self.numRegions + 1:
            raise ValueError("adjOffset must have numRegions + 1 entries")
        if self.r_xyz.shape[0] != 3 * self.numRegions:
            raise ValueError("r_xyz must have 3 * numRegions entries")
        self.numEdges = int(self.adjList.shape[0])
        self.device = device
        self._ctx = C.c_void_p()
        self._mesh = C.c_void_p()
        d = self.lib.dll
        self.lib.check(d.pb_context_create(device, C.byref(self._ctx)))
        self.lib.check(d.pb_mesh_create(self._ctx, self.numRegions, self.adjOffset.ctypes.data,
                                        self.adjList.ctypes.data, self.r_xyz.ctypes.data, C.byref(self._mesh)))
        self._mode = _lib.POINTER_HOST

    @classmethod
    def from_points(cls, r_xyz, device: int = 0, lib: Library | None = None, order: str = "canonical") -> "DeviceMesh":
        """buildSphere's triangulation on the device (js/sphere-mesh.js:174-186 → csrc/pb_meshgen.h): r_xyz holds the
        unit vectors of all regions, pole vertex included; adjOffset / adjList come back from the GPU.
        order="delaunator": the reference's own neighbour order (Delaunator 5.0.1's triangle numbering, host algorithm)."""
        self = cls.__new__(cls)
        self.lib = lib or default_library()
        self.device = device
        self._ctx = C.c_void_p()
        self._mesh = None
        self.lib.check(self.lib.dll.pb_context_create(device, C.byref(self._ctx)))
        return self._finish_from_points(r_xyz, order)

    @classmethod
    def build_sphere(cls, N: int, jitter: float, seed: float, device: int = 0, lib: Library | None = None,
                     order: str = "canonical") -> "DeviceMesh":
        """buildSphere(N, jitter, rng) (js/sphere-mesh.js:174-186) entirely on the device: Fibonacci points with
        makeRng(seed) jitter + the pole vertex, then the triangulation.  The result carries r_xyz, adjOffset, adjList.
        order="canonical" (default): device builder, every neighbour row starts at a canonical triangle;
        order="delaunator": rows start where the reference's SphereMesh constructor starts them (Delaunator 5.0.1's triangle
        numbering, a serial host algorithm) — same seed, same planet as the web app."""
        self = cls.__new__(cls)
        self.lib = lib or default_library()
        self.device = device
        self._ctx = C.c_void_p()
        self._mesh = None
        self.lib.check(self.lib.dll.pb_context_create(device, C.byref(self._ctx)))
        xyz = np.empty(3 * (int(N) + 1), np.float32)
        self.lib.check(self.lib.dll.pb_generate_fibonacci_sphere(self._ctx, int(N), float(jitter), float(seed), xyz.ctypes.data))
        return self._finish_from_points(xyz, order)

    @classmethod
    def _adopt(cls, parent: "DeviceMesh", mesh_handle, r_xyz) -> "DeviceMesh":
        """Wrap a pb_mesh created by the library on `parent`'s context (the coarse mesh of generateCoarsePlates)."""
        self = cls.__new__(cls)
        self.lib, self.device = parent.lib, parent.device
        self._ctx, self._owns_ctx, self._parent = parent._ctx, False, parent
        self._mesh = mesh_handle
        self.r_xyz = np.ascontiguousarray(r_xyz, np.float32).reshape(-1)
        self.numRegions = self.r_xyz.shape[0] // 3
        d = self.lib.dll
        self.numEdges = int(d.pb_mesh_num_edges(self._mesh))
        self.adjOffset = np.empty(self.numRegions + 1, np.int32)
        self.adjList = np.empty(self.numEdges, np.int32)
        self.lib.check(d.pb_mesh_get_adjacency(self._mesh, self.adjOffset.ctypes.data, self.adjList.ctypes.data))
        self._mode = parent._mode
        return self

    def _finish_from_points(self, r_xyz, order="canonical"):
        if order not in ("canonical", "delaunator"):
            raise ValueError("order must be 'canonical' or 'delaunator'")
        self.r_xyz = np.ascontiguousarray(r_xyz, np.float32).reshape(-1)
        self.numRegions = self.r_xyz.shape[0] // 3
        self._mesh = C.c_void_p()
        d = self.lib.dll
        # the option stays on the context: the coarse mesh of generateCoarsePlates follows the same order
        self.lib.check(d.pb_set_option(self._ctx, b"mesh_order", order.encode()))
        self.lib.check(d.pb_mesh_create_from_points(self._ctx, self.numRegions, self.r_xyz.ctypes.data, C.byref(self._mesh)))
        self.numEdges = int(d.pb_mesh_num_edges(self._mesh))
        self.adjOffset = np.empty(self.numRegions + 1, np.int32)
        self.adjList = np.empty(self.numEdges, np.int32)
        self.lib.check(d.pb_mesh_get_adjacency(self._mesh, self.adjOffset.ctypes.data, self.adjList.ctypes.data))
        self._mode = _lib.POINTER_HOST
        return self

    def close(self):
        st = getattr(self, "_climate", None)
        if st is not None:                # the climate state points into the mesh: release it first
            st.close()
            self._climate = None
        if getattr(self, "_mesh", None):
            self.lib.dll.pb_mesh_destroy(self._mesh)
            self._mesh = None
        if getattr(self, "_ctx", None) and getattr(self, "_owns_ctx", True):
            self.lib.dll.pb_context_destroy(self._ctx)
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- argument marshalling ---------------------------------------------------------------------
    def _begin(self, *arrays):
        """Pick the pointer mode from the argument types (all host or all device)."""
        present = [a for a in arrays if a is not None]
        dev = [_is_torch_cuda(a) for a in present]
        if any(dev) and not all(dev):
            raise TypeError("mix of host (numpy) and device (torch.cuda) arrays in one call")
        mode = _lib.POINTER_DEVICE if (dev and dev[0]) else _lib.POINTER_HOST
        if mode != self._mode:
            self.lib.check(self.lib.dll.pb_set_pointer_mode(self._ctx, mode))
            self._mode = mode
        if mode == _lib.POINTER_DEVICE:
            import torch
            self.lib.check(self.lib.dll.pb_set_stream(self._ctx, torch.cuda.current_stream().cuda_stream))
        return mode

    def _ptr(self, a, kind: str, n: int, name: str, optional: bool = False):
        if a is None:
            if optional:
                return None
            raise ValueError(f"{name} is required")
        if _is_torch_cuda(a):
            import torch
            want = {"f32": torch.float32, "i32": torch.int32, "u8": torch.uint8}[kind]
            if a.dtype != want or not a.is_contiguous() or a.numel() != n:
                raise ValueError(f"{name}: need a contiguous {want} CUDA tensor with {n} elements")
            return a.data_ptr()
        if not isinstance(a, np.ndarray) or a.dtype != _NP[kind] or not a.flags.c_contiguous or a.size != n:
            raise ValueError(f"{name}: need a C-contiguous numpy {_NP[kind].__name__} array with {n} elements")
        return a.ctypes.data

    def _new(self, like, kind: str, n: int):
        if _is_torch_cuda(like):
            import torch
            dt = {"f32": torch.float32, "i32": torch.int32, "u8": torch.uint8}[kind]
            return torch.empty(n, dtype=dt, device=like.device)
        return np.empty(n, _NP[kind])

    def set_option(self, name: str, value: str):
        """Engine options, e.g. ("flood", "host" | "device")."""
        self.lib.check(self.lib.dll.pb_set_option(self._ctx, name.encode(), value.encode()))

    def synchronize(self):
        self.lib.check(self.lib.dll.pb_synchronize(self._ctx))

    def profile_start(self, name_filter: str | None = None):
        f = name_filter.encode() if name_filter else None
        self.lib.check(self.lib.dll.pb_profile_start(self._ctx, f))

    def profile_stop(self) -> list:
        import json
        buf = C.create_string_buffer(1 << 16)
        self.lib.check(self.lib.dll.pb_profile_stop(self._ctx, buf, len(buf)))
        return json.loads(buf.value.decode())

    def launch_count(self) -> int:
        return int(self.lib.dll.pb_launch_count())

    # ---- mesh primitives ------------------------------------------------------------------------------
    @property
    def numTriangles(self) -> int:
        return 2 * self.numRegions - 4

    def trianglesAndHalfedges(self):
        """SphereMesh.triangles / .halfedges (js/sphere-mesh.js:94-100), canonical numbering, as numpy int32 arrays."""
        t = np.empty(3 * self.numTriangles, np.int32)
        h = np.empty(3 * self.numTriangles, np.int32)
        self._begin(t, h)
        self.lib.check(self.lib.dll.pb_mesh_get_triangles(self._mesh, t.ctypes.data, h.ctypes.data))
        return t, h

    def adjTriList(self):
        """SphereMesh._adjTriList (js/sphere-mesh.js:128-143): inner triangle of every adjacency slot (numpy int32[numEdges])."""
        out = np.empty(self.numEdges, np.int32)
        self._begin(out)
        self.lib.check(self.lib.dll.pb_mesh_get_adj_triangles(self._mesh, out.ctypes.data))
        return out

    def generateTriangleCenters(self, out=None):
        """generateTriangleCenters(mesh, r_xyz) (js/sphere-mesh.js:206-219)"""
        if out is None:
            out = np.empty(3 * self.numTriangles, np.float32)
        self._begin(out)
        self.lib.check(self.lib.dll.pb_generate_triangle_centers(self._mesh, self._ptr(out, "f32", 3 * self.numTriangles, "t_xyz")))
        return out

    def computeTriangleElevations(self, r_elevation, out=None):
        """computeTriangleElevations(mesh, r_elevation) (js/planet-worker.js:29-37)"""
        if out is None:
            out = self._new(r_elevation, "f32", self.numTriangles)
        self._begin(r_elevation, out)
        self.lib.check(self.lib.dll.pb_compute_triangle_elevations(
            self._mesh, self._ptr(r_elevation, "f32", self.numRegions, "r_elevation"), self._ptr(out, "f32", self.numTriangles, "t_elevation")))
        return out

    COLOR_MODES = {"terrain": 0, "biome": 1, "heightmap": 2, "landheightmap": 3, "landmask": 4, "biomeRaw": 5, "koppen": 6}

    def regionColors(self, mode: str, r_elevation, r_koppen=None, out=None):
        """Per-region r,g,b (Float32, 3·numRegions) of one of the renderer's colour modes: elevationToColor, smoothBiomeColors /
        biomeColor, heightmapColor, landHeightmapColor, landMaskColor, koppenColor (js/color-map.js:73-125, js/planet-mesh.js:30-80, 175-178)."""
        n = self.numRegions
        if out is None:
            out = self._new(r_elevation, "f32", 3 * n)
        self._begin(r_elevation, r_koppen, out)
        self.lib.check(self.lib.dll.pb_region_colors(
            self._mesh, self.COLOR_MODES[mode], self._ptr(r_elevation, "f32", n, "r_elevation"),
            None if r_koppen is None else self._ptr(r_koppen, "u8", n, "r_koppen"), self._ptr(out, "f32", 3 * n, "rgb")))
        return out

    def generateFibonacciSphere(self, N: int, jitter: float, seed: float, out=None):
        """generateFibonacciSphere + the pole vertex (js/sphere-mesh.js:9-37, 179-183): 3·(N+1) floats."""
        if out is None:
            out = np.empty(3 * (int(N) + 1), np.float32)
        self._begin(out)
        self.lib.check(self.lib.dll.pb_generate_fibonacci_sphere(self._ctx, int(N), C.c_double(jitter), C.c_double(seed),
                                                                  self._ptr(out, "f32", 3 * (int(N) + 1), "out")))
        return out

    def triangulateSphere(self, r_xyz, adjOffset=None, adjList=None):
        """Spherical Delaunay adjacency of `r_xyz` (buildSphere's triangulation + the SphereMesh constructor,
        js/sphere-mesh.js:94-146, 174-186) on this context's GPU.  numpy in → numpy out, torch.cuda in → torch.cuda out."""
        n = int(r_xyz.shape[0] if r_xyz.ndim == 1 else r_xyz.shape[0] * r_xyz.shape[1]) // 3
        if adjOffset is None:
            if _is_torch_cuda(r_xyz):
                import torch
                adjOffset = torch.empty(n + 1, dtype=torch.int32, device=r_xyz.device)
                adjList = torch.empty(6 * n - 12, dtype=torch.int32, device=r_xyz.device)
            else:
                adjOffset, adjList = np.empty(n + 1, np.int32), np.empty(6 * n - 12, np.int32)
        self._begin(r_xyz, adjOffset, adjList)
        self.lib.check(self.lib.dll.pb_triangulate_sphere(
            self._ctx, n, self._ptr(r_xyz, "f32", 3 * n, "r_xyz"), self._ptr(adjOffset, "i32", n + 1, "adjOffset"),
            self._ptr(adjList, "i32", 6 * n - 12, "adjList")))
        return adjOffset, adjList

    def computeNeighborDist(self, out=None):
        """js/sphere-mesh.js:191-203"""
        if out is None:
            out = np.empty(self.numEdges, np.float32)
        self._begin(out)
        self.lib.check(self.lib.dll.pb_compute_neighbor_dist(self._mesh, self._ptr(out, "f32", self.numEdges, "out")))
        return out
