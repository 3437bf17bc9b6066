"""TEST INFRASTRUCTURE ONLY.  Builds bindings/node/planet_b200_addon.cc together with the minimal in-process Node-API runtime
(napi_host.cc) against one of the two builds of the C ABI — the host emulation (CPU suite) or libplanet_b200.so (-m gpu) — and
drives the addon's exports from Python the way planet_worker_shim.mjs would from JavaScript: numbers, strings, plain objects
and typed arrays (numpy arrays are passed as views, so in-place stage functions mutate the caller's array like in JS)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# napi_typedarray_type (node_api.h order)
_NP_OF = {1: np.uint8, 2: np.uint8, 5: np.int32, 7: np.float32, 8: np.float64}
_TYPE_OF = {np.dtype(np.uint8): 1, np.dtype(np.int32): 5, np.dtype(np.float32): 7, np.dtype(np.float64): 8}
_UNDEFINED, _NULL, _BOOLEAN, _NUMBER, _STRING, _OBJECT, _FUNCTION = 0, 1, 2, 3, 4, 6, 7


class JsError(Exception):
    """`throw new Error(msg)` from the addon"""


class JsTypeError(JsError):
    """`throw new TypeError(msg)` from the addon"""


class Uint8Clamped(np.ndarray):
    """marker subclass: pass as a Uint8ClampedArray instead of a Uint8Array"""


def build(backend_so: str, tag: str, force: bool = False) -> str:
    so = os.path.join(HERE, f"libnapi_host_{tag}_TESTONLY.so")
    srcs = [os.path.join(HERE, "napi_host.cc"), os.path.join(ROOT, "bindings", "node", "planet_b200_addon.cc"),
            os.path.join(ROOT, "bindings", "node", "stub", "node_api.h"), os.path.join(ROOT, "include", "planet_b200.h"), backend_so]
    if not force and os.path.exists(so) and all(os.path.getmtime(s) <= os.path.getmtime(so) for s in srcs):
        return so
    d, f = os.path.split(backend_so)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-Werror", "-o", so, srcs[0], srcs[1],
                           "-I" + os.path.join(ROOT, "bindings", "node", "stub"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + d, "-l:" + f, "-Wl,-rpath," + d])
    return so


class NapiHost:
    def __init__(self, backend_so: str, tag: str):
        self.dll = d = C.CDLL(build(backend_so, tag))
        vp = C.c_void_p
        for name, res, args in [
                ("nh_init", C.c_int, []), ("nh_release", None, []), ("nh_undefined", vp, []), ("nh_null", vp, []),
                ("nh_number", vp, [C.c_double]), ("nh_bool", vp, [C.c_int]), ("nh_string", vp, [C.c_char_p]), ("nh_object", vp, []),
                ("nh_typed_view", vp, [C.c_int, C.c_size_t, vp]), ("nh_set", C.c_int, [vp, C.c_char_p, vp]),
                ("nh_get", vp, [vp, C.c_char_p]), ("nh_typeof", C.c_int, [vp]), ("nh_number_value", C.c_double, [vp]),
                ("nh_bool_value", C.c_int, [vp]), ("nh_string_value", C.c_char_p, [vp]), ("nh_is_typed", C.c_int, [vp]),
                ("nh_typed_info", C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(vp)]),
                ("nh_num_keys", C.c_size_t, [vp]), ("nh_key", C.c_char_p, [vp, C.c_size_t]), ("nh_num_exports", C.c_size_t, []),
                ("nh_export_name", C.c_char_p, [C.c_size_t]), ("nh_call", vp, [C.c_char_p, C.c_int, C.POINTER(vp)]),
                ("nh_take_exception", C.c_int, [C.c_char_p, C.c_size_t])]:
            fn = getattr(d, name)
            fn.restype, fn.argtypes = res, args
        if d.nh_init() != 0:
            raise RuntimeError("addon registration failed: " + self._exception()[1])

    @property
    def exports(self) -> list:
        return [self.dll.nh_export_name(i).decode() for i in range(self.dll.nh_num_exports())]

    def _exception(self):
        buf = C.create_string_buffer(1024)
        kind = self.dll.nh_take_exception(buf, len(buf))
        return kind, buf.value.decode("utf-8", "replace")

    def _to_js(self, v, keep):
        d = self.dll
        if v is None:
            return d.nh_null()
        if isinstance(v, bool):
            return d.nh_bool(int(v))
        if isinstance(v, (int, float, np.integer, np.floating)):
            return d.nh_number(float(v))
        if isinstance(v, str):
            return d.nh_string(v.encode())
        if isinstance(v, np.ndarray):
            if not v.flags.c_contiguous or v.dtype not in _TYPE_OF:
                raise TypeError("typed-array arguments must be C-contiguous uint8 / int32 / float32 / float64 numpy arrays")
            keep.append(v)
            t = 2 if isinstance(v, Uint8Clamped) else _TYPE_OF[v.dtype]
            return d.nh_typed_view(t, v.size, v.ctypes.data)
        if isinstance(v, dict):
            o = d.nh_object()
            for k, x in v.items():
                d.nh_set(o, k.encode(), self._to_js(x, keep))
            return o
        raise TypeError(f"cannot pass {type(v).__name__} to the addon")

    def _from_js(self, h):
        d = self.dll
        t = d.nh_typeof(h)
        if t in (_UNDEFINED, _NULL):
            return None
        if t == _BOOLEAN:
            return bool(d.nh_bool_value(h))
        if t == _NUMBER:
            return d.nh_number_value(h)
        if t == _STRING:
            return d.nh_string_value(h).decode()
        if t == _OBJECT and d.nh_is_typed(h):
            ty, n, p = C.c_int(), C.c_size_t(), C.c_void_p()
            d.nh_typed_info(h, C.byref(ty), C.byref(n), C.byref(p))
            dt = np.dtype(_NP_OF[ty.value])
            out = np.empty(n.value, dt)
            if n.value:
                C.memmove(out.ctypes.data, p.value, n.value * dt.itemsize)
            return out.view(Uint8Clamped) if ty.value == 2 else out
        if t == _OBJECT:
            return {d.nh_key(h, i).decode(): self._from_js(d.nh_get(h, d.nh_key(h, i))) for i in range(d.nh_num_keys(h))}
        raise TypeError(f"unsupported JS value type {t}")

    def call(self, name: str, *args):
        keep = []
        argv = (C.c_void_p * max(1, len(args)))(*[self._to_js(a, keep) for a in args])
        r = self.dll.nh_call(name.encode(), len(args), argv)
        try:
            if not r:
                kind, msg = self._exception()
                raise (JsTypeError if kind == 2 else JsError)(msg)
            return self._from_js(r)
        finally:
            self.dll.nh_release()

    def __getattr__(self, name):
        if name.startswith("_") or name in ("dll",):
            raise AttributeError(name)
        return lambda *a: self.call(name, *a)
