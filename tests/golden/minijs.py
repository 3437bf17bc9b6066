"""TEST INFRASTRUCTURE ONLY — a small evaluator for the ECMAScript subset parsed by minijs_parser.py.

It exists to EXECUTE the unmodified reference (/root/reference/js/*.js: rng, simplex-noise, sphere-mesh, plates, ocean-land,
coarse-plates, super-plates, elevation, terrain-post, climate-util, color-map, wind, ocean, heuristic-precip, precipitation,
temperature, koppen and the worker planet-worker.js) in the build container, where no JavaScript runtime exists, so that golden
vectors produced by the reference's own source pin the oracle (tests/golden/make_reference_vectors.py).  It is not a JavaScript
engine: the language subset is what those files use, and anything else raises.

Semantics that matter for numeric parity and are implemented faithfully:
  * every number is an IEEE double (Python float); `|0`, `>>>`, `^` … go through ToInt32 / ToUint32 on the double
  * typed arrays round on store (Float32Array: round-to-nearest-even via a C float cast; integer arrays wrap modulo 2^k),
    `subarray` shares memory, reads outside the array give `undefined`
  * `%` keeps the dividend's sign, division by zero gives ±Infinity / NaN, `Math.round` is floor(x + 0.5),
    `Math.max / min` propagate NaN and order -0 < +0
  * plain objects enumerate integer-like keys in ascending order first, then string keys in insertion order;
    `Set` / `Map` iterate in insertion order and see entries added during the iteration; `Array.prototype.sort` is stable
  * `let`/`const` block scopes, closures, classes with getters, arrow functions with lexical `this`
Math.sin/cos/exp/pow/log/atan2/asin/tanh are Python's libm (glibc), i.e. a third implementation next to V8's fdlibm port and
include/pb_detmath.h: all three agree to within an ulp or two of a double, which the Float32 stores hide except with a
probability of order 1e-8 per evaluation.
"""
from __future__ import annotations

import array
import functools
import math
import os
import time

from .minijs_parser import parse

# ---------------------------------------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------------------------------------


class _Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = _Undefined()
NAN = float("nan")
INF = float("inf")


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(to_display(value))
        self.value = value


class JSObject:
    __slots__ = ("props", "proto", "cls")

    def __init__(self, proto=None, props=None):
        self.props = {} if props is None else props
        self.proto = proto
        self.cls = None


class JSProxy:
    """new Proxy(target, {get, has, ownKeys}) — the three traps bindings/node/planet_worker_shim.mjs uses"""
    __slots__ = ("target", "handler")

    def __init__(self, target, handler):
        self.target, self.handler = target, handler

    def trap(self, name):
        f = self.handler.props.get(name)
        return f if f is not None and f is not UNDEF else None


class Accessor:
    __slots__ = ("get", "set")

    def __init__(self, get=None, set=None):
        self.get, self.set = get, set


class JSArray(list):
    __slots__ = ()


class JSFunction:
    __slots__ = ("name", "nparams", "bind_params", "body", "env", "is_arrow", "expr_body", "this_val", "props", "fields", "is_class", "line",
                 "fname", "var_names")

    def __init__(self):
        self.props = {}
        self.fields = None
        self.is_class = False
        self.this_val = UNDEF

    def call(self, this, args):
        env = Env({}, self.env)
        v = env.v
        if self.is_arrow:
            this = self.this_val
        v["this"] = this
        for name in self.var_names:
            v[name] = UNDEF
        self.bind_params(env, args)
        if self.expr_body:
            return self.body(env)
        r = self.body(env)
        if r is not None and r.__class__ is Ret:
            return r.value
        return UNDEF


class HostFunction:
    __slots__ = ("fn", "name", "construct", "props")

    def __init__(self, fn, name="", construct=None, props=None):
        self.fn, self.name, self.construct = fn, name, construct
        self.props = props or {}


class Env:
    __slots__ = ("v", "p")

    def __init__(self, v, p):
        self.v, self.p = v, p


class Ret:
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class _Signal:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name


BRK, CNT = _Signal("break"), _Signal("continue")

_TYPECODES = {"Float32Array": "f", "Float64Array": "d", "Int32Array": "i", "Uint32Array": "I", "Int16Array": "h", "Uint16Array": "H",
              "Int8Array": "b", "Uint8Array": "B", "Uint8ClampedArray": "B"}
_INT_BITS = {"i": 32, "I": 32, "h": 16, "H": 16, "b": 8, "B": 8}
_SIGNED = {"i", "h", "b"}


class ArrayBuffer:
    __slots__ = ("arr",)

    def __init__(self, arr):
        self.arr = arr


class JSTypedArray:
    """view [off, off+length) of an array.array; subarray() shares the storage"""
    __slots__ = ("kind", "code", "mv", "length", "base", "clamped")

    def __init__(self, kind, n=0, base=None, mv=None):
        self.kind = kind
        self.code = _TYPECODES[kind]
        self.clamped = kind == "Uint8ClampedArray"
        if mv is None:
            base = array.array(self.code, bytes(array.array(self.code).itemsize * n))
            mv = memoryview(base)
        self.base, self.mv, self.length = base, mv, len(mv)

    def get(self, i):
        if 0 <= i < self.length:
            return float(self.mv[i])
        return UNDEF

    def store(self, i, v):
        if not (0 <= i < self.length):
            return
        c = self.code
        if c == "f" or c == "d":
            x = v if type(v) is float else to_num(v)
            try:
                self.mv[i] = x
            except OverflowError:
                self.mv[i] = INF if x > 0 else -INF
        else:
            self.mv[i] = self.convert_int(v)

    def convert_int(self, v):
        x = v if type(v) is float else to_num(v)
        if x != x or x in (INF, -INF):
            return 0
        if self.clamped:
            return 0 if x < 0 else 255 if x > 255 else int(round_half_even(x))
        bits = _INT_BITS[self.code]
        n = int(x) & ((1 << bits) - 1)
        if self.code in _SIGNED and n >= 1 << (bits - 1):
            n -= 1 << bits
        return n

    def tolist(self):
        return self.mv.tolist()


def round_half_even(x):
    return round(x)


class JSSet:
    __slots__ = ("items", "index", "live")

    def __init__(self):
        self.items, self.index, self.live = [], {}, 0

    @staticmethod
    def key(v):
        if type(v) is float:
            if v != v:
                return ("nan",)
            return v + 0.0 if v != 0 else 0.0
        if type(v) is bool:
            return ("bool", v)
        if isinstance(v, (str, type(None), _Undefined)):
            return v if isinstance(v, str) else ("nullish", v is None)
        return ("obj", id(v))

    def add(self, v):
        k = self.key(v)
        if k not in self.index:
            self.index[k] = len(self.items)
            self.items.append((True, v))
            self.live += 1

    def has(self, v):
        return self.key(v) in self.index

    def delete(self, v):
        k = self.key(v)
        i = self.index.pop(k, None)
        if i is None:
            return False
        self.items[i] = (False, None)
        self.live -= 1
        return True

    def clear(self):
        for i in range(len(self.items)):
            self.items[i] = (False, None)
        self.index.clear()
        self.live = 0

    def iterate(self):
        i = 0
        items = self.items
        while i < len(items):
            ok, v = items[i]
            if ok:
                yield v
            i += 1


class JSMap:
    __slots__ = ("items", "index", "live")

    def __init__(self):
        self.items, self.index, self.live = [], {}, 0

    def set(self, k, v):
        kk = JSSet.key(k)
        i = self.index.get(kk)
        if i is None:
            self.index[kk] = len(self.items)
            self.items.append([True, k, v])
            self.live += 1
        else:
            self.items[i][2] = v

    def get(self, k):
        i = self.index.get(JSSet.key(k))
        return UNDEF if i is None else self.items[i][2]

    def has(self, k):
        return JSSet.key(k) in self.index

    def delete(self, k):
        i = self.index.pop(JSSet.key(k), None)
        if i is None:
            return False
        self.items[i] = [False, None, None]
        self.live -= 1
        return True

    def clear(self):
        for it in self.items:
            it[0] = False
        self.index.clear()
        self.live = 0

    def iterate(self):
        i = 0
        items = self.items
        while i < len(items):
            ok, k, v = items[i]
            if ok:
                yield k, v
            i += 1


# ---------------------------------------------------------------------------------------------------------------------
# conversions and operators
# ---------------------------------------------------------------------------------------------------------------------


def number_to_string(x: float) -> str:
    """Number::toString(10)"""
    if x != x:
        return "NaN"
    if x == 0:
        return "0"
    if x in (INF, -INF):
        return "Infinity" if x > 0 else "-Infinity"
    if x < 0:
        return "-" + number_to_string(-x)
    r = repr(x)
    if "e" in r:
        mant, e = r.split("e")
        exp10 = int(e)
    else:
        mant, exp10 = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    if fp == "0":
        fp = ""
    digits = (ip + fp).lstrip("0")
    # value = 0.digits × 10^n
    n = len(ip.lstrip("0")) + exp10 if ip.strip("0") else exp10 - (len(fp) - len(fp.lstrip("0")))
    digits = digits.rstrip("0") or "0"
    k = len(digits)
    if k <= n <= 21:
        return digits + "0" * (n - k)
    if 0 < n <= 21:
        return digits[:n] + "." + digits[n:]
    if -6 < n <= 0:
        return "0." + "0" * (-n) + digits
    e = n - 1
    sign = "+" if e >= 0 else "-"
    if k == 1:
        return f"{digits}e{sign}{abs(e)}"
    return f"{digits[0]}.{digits[1:]}e{sign}{abs(e)}"


def to_str(v) -> str:
    t = type(v)
    if t is str:
        return v
    if t is float:
        return number_to_string(v)
    if t is bool:
        return "true" if v else "false"
    if v is None:
        return "null"
    if v is UNDEF:
        return "undefined"
    if t is JSArray:
        return ",".join("" if (x is None or x is UNDEF) else to_str(x) for x in v)
    if t is JSTypedArray:
        return ",".join(to_str(float(x)) for x in v.mv)
    if t is JSObject:
        if "message" in v.props and v.cls == "Error":
            return f"{to_str(v.props.get('name', 'Error'))}: {to_str(v.props['message'])}"
        return "[object Object]"
    if t in (JSFunction, HostFunction):
        return f"function {v.name}() {{ … }}"
    return str(v)


def to_display(v):
    try:
        return to_str(v)
    except Exception:
        return repr(v)


def to_num(v) -> float:
    t = type(v)
    if t is float:
        return v
    if t is bool:
        return 1.0 if v else 0.0
    if v is None:
        return 0.0
    if v is UNDEF:
        return NAN
    if t is str:
        s = v.strip()
        if not s:
            return 0.0
        try:
            if s[:2] in ("0x", "0X"):
                return float(int(s[2:], 16))
            if s in ("Infinity", "+Infinity"):
                return INF
            if s == "-Infinity":
                return -INF
            if any(c in s for c in "_nN") or s.lower() in ("inf", "-inf", "+inf"):
                return NAN
            return float(s)
        except ValueError:
            return NAN
    if t is JSArray:
        if len(v) == 0:
            return 0.0
        if len(v) == 1:
            return to_num(to_str(v))
        return NAN
    return NAN


def truthy(v) -> bool:
    t = type(v)
    if t is bool:
        return v
    if t is float:
        return v == v and v != 0.0
    if t is str:
        return len(v) > 0
    if v is None or v is UNDEF:
        return False
    return True


def to_int32(v) -> int:
    x = v if type(v) is float else to_num(v)
    if x != x or x == INF or x == -INF:
        return 0
    n = int(x) & 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


def to_uint32(v) -> int:
    x = v if type(v) is float else to_num(v)
    if x != x or x == INF or x == -INF:
        return 0
    return int(x) & 0xFFFFFFFF


def js_div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return NAN
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -INF if neg else INF


def js_mod(a, b):
    try:
        return math.fmod(a, b)
    except ValueError:
        return NAN


def js_pow(a, b):
    if b != b:
        return NAN
    if b == 0:
        return 1.0
    if (a == 1 or a == -1) and b in (INF, -INF):
        return NAN
    try:
        return math.pow(a, b)
    except OverflowError:
        if a < 0 and b == math.floor(b) and math.fmod(b, 2.0) != 0:
            return -INF
        return INF
    except ValueError:
        if a == 0:
            # 0 ** negative
            if math.copysign(1.0, a) < 0 and b == math.floor(b) and math.fmod(b, 2.0) != 0:
                return -INF
            return INF
        return NAN
    except ZeroDivisionError:
        return INF


def typeof(v) -> str:
    t = type(v)
    if t is float:
        return "number"
    if t is str:
        return "string"
    if t is bool:
        return "boolean"
    if v is UNDEF:
        return "undefined"
    if t in (JSFunction, HostFunction):
        return "function"
    return "object"


def strict_equals(a, b) -> bool:
    ta, tb = type(a), type(b)
    if ta is float and tb is float:
        return a == b
    if ta is not tb:
        return False
    if ta is str or ta is bool:
        return a == b
    return a is b


def loose_equals(a, b) -> bool:
    if (a is None or a is UNDEF) and (b is None or b is UNDEF):
        return True
    if a is None or a is UNDEF or b is None or b is UNDEF:
        return False
    ta, tb = type(a), type(b)
    if ta is tb:
        return strict_equals(a, b)
    prim = (float, str, bool)
    if ta in prim and tb in prim:
        return to_num(a) == to_num(b)
    return False


def js_add(a, b):
    ta, tb = type(a), type(b)
    if ta is float and tb is float:
        return a + b
    if ta is str or tb is str or ta in (JSArray, JSObject, JSTypedArray) or tb in (JSArray, JSObject, JSTypedArray):
        return to_str(a) + to_str(b)
    return to_num(a) + to_num(b)


def js_less(a, b, swap=False, orequal=False):
    """a < b (orequal: a <= b) with string / number semantics"""
    if type(a) is str and type(b) is str:
        return (a <= b) if orequal else (a < b)
    x, y = to_num(a), to_num(b)
    return (x <= y) if orequal else (x < y)


def throw_error(kind, msg):
    raise JSThrow(make_error(kind, msg))


def make_error(kind, msg):
    o = JSObject(None, {"name": kind, "message": msg, "stack": f"{kind}: {msg}"})
    o.cls = "Error"
    return o


# ---------------------------------------------------------------------------------------------------------------------
# property access
# ---------------------------------------------------------------------------------------------------------------------


def is_array_index(k: str) -> bool:
    return k.isdigit() and (k == "0" or k[0] != "0") and len(k) < 10


def own_keys(o: JSObject):
    keys = list(o.props.keys())
    ints = [k for k in keys if is_array_index(k)]
    if not ints:
        return keys
    ints.sort(key=int)
    return ints + [k for k in keys if not is_array_index(k)]


def prop_key(k) -> str:
    return k if type(k) is str else to_str(k)


def index_of(k):
    """array index of a property key, or -1"""
    if type(k) is float:
        i = int(k) if k == k and k not in (INF, -INF) else -1
        return i if i == k and i >= 0 else -1
    if type(k) is str and is_array_index(k):
        return int(k)
    return -1


def bound(obj, fn):
    return HostFunction(lambda this, args: fn(obj, args), getattr(fn, "__name__", "method"))


def get_prop(obj, key, line=None):
    t = type(obj)
    if t is JSTypedArray:
        if type(key) is float:
            i = int(key)
            if i == key and 0 <= i < obj.length:
                return float(obj.mv[i])
            return UNDEF
        if key == "length":
            return float(obj.length)
        m = TYPED_METHODS.get(key)
        if m is not None:
            return bound(obj, m)
        if key == "buffer":
            return ArrayBuffer(obj.base)
        if key == "byteLength":
            return float(obj.length * obj.mv.itemsize)
        if key == "BYTES_PER_ELEMENT":
            return float(obj.mv.itemsize)
        i = index_of(key)
        return obj.get(i) if i >= 0 else UNDEF
    if t is JSArray:
        if type(key) is float:
            i = int(key)
            if i == key and 0 <= i < len(obj):
                v = obj[i]
                return v
            return UNDEF
        if key == "length":
            return float(len(obj))
        m = ARRAY_METHODS.get(key)
        if m is not None:
            return bound(obj, m)
        i = index_of(key)
        return obj[i] if 0 <= i < len(obj) else UNDEF
    if t is JSObject:
        k = key if type(key) is str else to_str(key)
        o = obj
        while o is not None:
            v = o.props.get(k, o)
            if v is not o:
                if type(v) is Accessor:
                    return call_function(v.get, obj, []) if v.get else UNDEF
                return v
            o = o.proto
        if k == "hasOwnProperty":
            return HostFunction(lambda this, args: prop_key(args[0]) in obj.props, "hasOwnProperty")
        if k == "toString":
            return HostFunction(lambda this, args: to_str(obj), "toString")
        return UNDEF
    if t is str:
        if type(key) is float:
            i = int(key)
            return obj[i] if i == key and 0 <= i < len(obj) else UNDEF
        if key == "length":
            return float(len(obj))
        m = STRING_METHODS.get(key)
        if m is not None:
            return bound(obj, m)
        return UNDEF
    if t is JSSet:
        if key == "size":
            return float(obj.live)
        m = SET_METHODS.get(key)
        return bound(obj, m) if m else UNDEF
    if t is JSMap:
        if key == "size":
            return float(obj.live)
        m = MAP_METHODS.get(key)
        return bound(obj, m) if m else UNDEF
    if t is float:
        m = NUMBER_METHODS.get(key)
        return bound(obj, m) if m else UNDEF
    if t is JSFunction or t is HostFunction:
        k = prop_key(key)
        if k in obj.props:
            return obj.props[k]
        if k == "name":
            return obj.name or ""
        if k == "call":
            return HostFunction(lambda this, args: call_function(obj, args[0] if args else UNDEF, list(args[1:])), "call")
        if k == "apply":
            return HostFunction(lambda this, args: call_function(obj, args[0] if args else UNDEF, iterate(args[1]) if len(args) > 1 and args[1] not in (None, UNDEF) else []), "apply")
        if k == "bind":
            return HostFunction(lambda this, args: HostFunction(lambda t2, a2: call_function(obj, args[0] if args else UNDEF, list(args[1:]) + list(a2)), "bound"), "bind")
        return UNDEF
    if t is JSProxy:
        f = obj.trap("get")
        k = key if type(key) is str else to_str(key)
        return call_function(f, obj.handler, [obj.target, k, obj]) if f else get_prop(obj.target, k)
    if t is _PyIter:
        if key == "next":
            def nxt(this, args):
                for v in obj.it:
                    return JSObject(None, {"value": v, "done": False})
                return JSObject(None, {"value": UNDEF, "done": True})
            return HostFunction(nxt, "next")
        return UNDEF
    if t is ArrayBuffer:
        if key == "byteLength":
            return float(len(obj.arr) * obj.arr.itemsize)
        return UNDEF
    if t is bool:
        return UNDEF
    where = f" (line {line})" if line else ""
    throw_error("TypeError", f"Cannot read properties of {to_display(obj)} (reading '{to_display(key)}'){where}")


def set_prop(obj, key, value):
    t = type(obj)
    if t is JSTypedArray:
        if type(key) is float:
            i = int(key)
            if i == key:
                obj.store(i, value)
            return
        i = index_of(key)
        if i >= 0:
            obj.store(i, value)
        return
    if t is JSArray:
        if type(key) is float:
            i = int(key)
            if i != key or i < 0:
                throw_error("TypeError", "non-index array properties are not supported")
        elif key == "length":
            n = int(to_num(value))
            if n < len(obj):
                del obj[n:]
            else:
                obj.extend([UNDEF] * (n - len(obj)))
            return
        else:
            i = index_of(key)
            if i < 0:
                throw_error("TypeError", f"array property {key!r} is not supported")
        if i < len(obj):
            obj[i] = value
        else:
            obj.extend([UNDEF] * (i - len(obj)))
            obj.append(value)
        return
    if t is JSObject:
        k = key if type(key) is str else to_str(key)
        o = obj.proto
        while o is not None:          # setter on the prototype chain
            v = o.props.get(k)
            if type(v) is Accessor:
                if v.set:
                    call_function(v.set, obj, [value])
                return
            o = o.proto
        obj.props[k] = value
        return
    if t is JSFunction or t is HostFunction:
        obj.props[prop_key(key)] = value
        return
    if t is JSProxy:
        set_prop(obj.target, key, value)
        return
    throw_error("TypeError", f"Cannot set properties of {to_display(obj)} (setting '{to_display(key)}')")


def call_function(f, this, args, line=None):
    t = type(f)
    if t is JSFunction:
        if f.is_class:
            throw_error("TypeError", f"Class constructor {f.name} cannot be invoked without 'new'")
        return f.call(this, args)
    if t is HostFunction:
        r = f.fn(this, args)
        return _wrap(r)
    where = f" (line {line})" if line else ""
    throw_error("TypeError", f"{to_display(f)} is not a function{where}")


def _wrap(r):
    """host results → JS values"""
    t = type(r)
    if t is int:
        return float(r)
    if r is None and t is type(None):
        return None
    return r


def construct(f, args, line=None):
    t = type(f)
    if t is JSFunction:
        if f.is_arrow:
            throw_error("TypeError", "arrow functions are not constructors")
        proto = f.props.get("prototype")
        obj = JSObject(proto)
        if f.fields:
            for k, init in f.fields:
                obj.props[k] = init(Env({"this": obj}, f.env)) if init else UNDEF
        saved, f.is_class = f.is_class, False
        try:
            r = f.call(obj, args)
        finally:
            f.is_class = saved
        return r if type(r) in (JSObject, JSArray, JSTypedArray, JSSet, JSMap) else obj
    if t is HostFunction and f.construct is not None:
        return f.construct(args)
    where = f" (line {line})" if line else ""
    throw_error("TypeError", f"{to_display(f)} is not a constructor{where}")


def iterate(v):
    """values of a JS iterable, live (growth during the iteration is seen)"""
    t = type(v)
    if t is JSArray:
        i = 0
        while i < len(v):
            yield v[i]
            i += 1
        return
    if t is JSTypedArray:
        i = 0
        while i < v.length:
            yield float(v.mv[i])
            i += 1
        return
    if t is JSSet:
        yield from v.iterate()
        return
    if t is JSMap:
        for k, x in v.iterate():
            yield JSArray([k, x])
        return
    if t is str:
        yield from v
        return
    if type(v) is _PyIter:
        yield from v.it
        return
    throw_error("TypeError", f"{to_display(v)} is not iterable")


class _PyIter:
    """iterator objects returned by .keys() / .values() / .entries()"""
    __slots__ = ("it",)

    def __init__(self, it):
        self.it = iter(it)


# ---------------------------------------------------------------------------------------------------------------------
# builtins
# ---------------------------------------------------------------------------------------------------------------------


def _arg(args, i, default=UNDEF):
    return args[i] if i < len(args) else default


def _rel_index(v, n, default):
    if v is UNDEF:
        return default
    x = to_num(v)
    if x != x:
        return 0
    x = math.trunc(x) if x not in (INF, -INF) else (n if x > 0 else -n)
    if x < 0:
        return max(0, n + int(x))
    return min(n, int(x))


def _cmp_key(cmpfn):
    def cmp(a, b):
        r = to_num(call_function(cmpfn, UNDEF, [a, b]))
        return -1 if r < 0 else 1 if r > 0 else 0
    return functools.cmp_to_key(cmp)


def _array_sort(a, args):
    cmpfn = _arg(args, 0)
    undef = [x for x in a if x is UNDEF]
    vals = [x for x in a if x is not UNDEF]
    if cmpfn is UNDEF:
        vals.sort(key=to_str)
    else:
        vals.sort(key=_cmp_key(cmpfn))
    a[:] = vals + undef
    return a


def _array_splice(a, args):
    n = len(a)
    start = _rel_index(_arg(args, 0), n, 0)
    count = n - start if len(args) < 2 else max(0, min(n - start, int(to_num(args[1]))))
    removed = JSArray(a[start:start + count])
    a[start:start + count] = list(args[2:])
    return removed


def _flat(a, depth):
    out = JSArray()
    for x in a:
        if type(x) is JSArray and depth > 0:
            out.extend(_flat(x, depth - 1))
        else:
            out.append(x)
    return out


def _index_of_value(seq, v, start=0):
    for i in range(start, len(seq)):
        if strict_equals(seq[i], v):
            return float(i)
    return -1.0


def _includes(seq, v):
    for x in seq:
        if strict_equals(x, v) or (type(v) is float and v != v and type(x) is float and x != x):
            return True
    return False


def _reduce(a, args):
    it = list(a)
    i = 0
    if len(args) > 1:
        acc = args[1]
    else:
        if not it:
            throw_error("TypeError", "Reduce of empty array with no initial value")
        acc, i = it[0], 1
    while i < len(it):
        acc = call_function(args[0], UNDEF, [acc, it[i], float(i), a])
        i += 1
    return acc


ARRAY_METHODS = {
    "push": lambda a, args: (a.extend(args), float(len(a)))[1],
    "pop": lambda a, args: a.pop() if a else UNDEF,
    "shift": lambda a, args: a.pop(0) if a else UNDEF,
    "unshift": lambda a, args: (a.__setitem__(slice(0, 0), list(args)), float(len(a)))[1],
    "slice": lambda a, args: JSArray(a[_rel_index(_arg(args, 0), len(a), 0):_rel_index(_arg(args, 1), len(a), len(a))]),
    "splice": _array_splice,
    "concat": lambda a, args: JSArray(list(a) + [y for x in args for y in (x if type(x) is JSArray else [x])]),
    "indexOf": lambda a, args: _index_of_value(a, _arg(args, 0), _rel_index(_arg(args, 1), len(a), 0)),
    "includes": lambda a, args: _includes(a, _arg(args, 0)),
    "join": lambda a, args: (", " if False else (to_str(_arg(args, 0)) if _arg(args, 0) is not UNDEF else ",")).join(
        "" if (x is None or x is UNDEF) else to_str(x) for x in a),
    "reverse": lambda a, args: (a.reverse(), a)[1],
    "sort": _array_sort,
    "fill": lambda a, args: (a.__setitem__(slice(_rel_index(_arg(args, 1), len(a), 0), _rel_index(_arg(args, 2), len(a), len(a))),
                                           [_arg(args, 0)] * max(0, _rel_index(_arg(args, 2), len(a), len(a)) - _rel_index(_arg(args, 1), len(a), 0))), a)[1],
    "map": lambda a, args: JSArray([call_function(args[0], UNDEF, [x, float(i), a]) for i, x in enumerate(list(a))]),
    "filter": lambda a, args: JSArray([x for i, x in enumerate(list(a)) if truthy(call_function(args[0], UNDEF, [x, float(i), a]))]),
    "forEach": lambda a, args: ([call_function(args[0], UNDEF, [x, float(i), a]) for i, x in enumerate(list(a))], UNDEF)[1],
    "some": lambda a, args: any(truthy(call_function(args[0], UNDEF, [x, float(i), a])) for i, x in enumerate(list(a))),
    "every": lambda a, args: all(truthy(call_function(args[0], UNDEF, [x, float(i), a])) for i, x in enumerate(list(a))),
    "find": lambda a, args: next((x for i, x in enumerate(list(a)) if truthy(call_function(args[0], UNDEF, [x, float(i), a]))), UNDEF),
    "findIndex": lambda a, args: next((float(i) for i, x in enumerate(list(a)) if truthy(call_function(args[0], UNDEF, [x, float(i), a]))), -1.0),
    "reduce": _reduce,
    "flat": lambda a, args: _flat(a, int(to_num(_arg(args, 0, 1.0)))),
    "flatMap": lambda a, args: _flat(JSArray([call_function(args[0], UNDEF, [x, float(i), a]) for i, x in enumerate(list(a))]), 1),
    "keys": lambda a, args: _PyIter(float(i) for i in range(len(a))),
    "values": lambda a, args: _PyIter(iterate(a)),
    "entries": lambda a, args: _PyIter(JSArray([float(i), x]) for i, x in enumerate(list(a))),
}


def _typed_set(t, args):
    src, off = _arg(args, 0), int(to_num(_arg(args, 1, 0.0)))
    if type(src) is JSTypedArray:
        n = src.length
        if off + n > t.length:
            throw_error("RangeError", "offset is out of bounds")
        if src.code == t.code:
            t.mv[off:off + n] = src.mv          # memoryview handles overlap like memmove
        else:
            for i in range(n):
                t.store(off + i, float(src.mv[i]))
        return UNDEF
    vals = list(iterate(src))
    if off + len(vals) > t.length:
        throw_error("RangeError", "offset is out of bounds")
    for i, v in enumerate(vals):
        t.store(off + i, v)
    return UNDEF


def _typed_fill(t, args):
    b, e = _rel_index(_arg(args, 1), t.length, 0), _rel_index(_arg(args, 2), t.length, t.length)
    if e > b:
        if t.code in ("f", "d"):
            t.store(b, _arg(args, 0))
            v = t.mv[b]
        else:
            v = t.convert_int(_arg(args, 0))
        t.mv[b:e] = array.array(t.code, [v]) * (e - b)
    return t


def _typed_sub(t, args):
    b, e = _rel_index(_arg(args, 0), t.length, 0), _rel_index(_arg(args, 1), t.length, t.length)
    return JSTypedArray(t.kind, base=t.base, mv=t.mv[b:max(b, e)])


def _typed_slice(t, args):
    b, e = _rel_index(_arg(args, 0), t.length, 0), _rel_index(_arg(args, 1), t.length, t.length)
    base = array.array(t.code, t.mv[b:max(b, e)].tobytes())
    return JSTypedArray(t.kind, base=base, mv=memoryview(base))


def _typed_sort(t, args):
    vals = t.mv.tolist()
    cmpfn = _arg(args, 0)
    if cmpfn is UNDEF:
        nans = [x for x in vals if x != x]
        vals = [x for x in vals if x == x]
        vals.sort(key=lambda x: (x, 0 if math.copysign(1.0, x) < 0 else 1) if x == 0 else (x, 0))
        vals += nans
    else:
        vals = [float(x) for x in vals]
        vals.sort(key=_cmp_key(cmpfn))
    t.mv[:] = array.array(t.code, [int(x) for x in vals] if t.code not in ("f", "d") else vals)
    return t


def _typed_from_iter(kind, values):
    vals = list(values)
    t = JSTypedArray(kind, len(vals))
    for i, v in enumerate(vals):
        t.store(i, v)
    return t


TYPED_METHODS = {
    "set": _typed_set, "fill": _typed_fill, "subarray": _typed_sub, "slice": _typed_slice, "sort": _typed_sort,
    "indexOf": lambda t, args: _index_of_value([float(x) for x in t.mv], _arg(args, 0)),
    "includes": lambda t, args: _includes([float(x) for x in t.mv], _arg(args, 0)),
    "join": lambda t, args: (to_str(_arg(args, 0)) if _arg(args, 0) is not UNDEF else ",").join(to_str(float(x)) for x in t.mv),
    "map": lambda t, args: _typed_from_iter(t.kind, [call_function(args[0], UNDEF, [float(x), float(i), t]) for i, x in enumerate(t.mv.tolist())]),
    "forEach": lambda t, args: ([call_function(args[0], UNDEF, [float(x), float(i), t]) for i, x in enumerate(t.mv.tolist())], UNDEF)[1],
    "reduce": lambda t, args: _reduce(JSArray(float(x) for x in t.mv), args),
    "some": lambda t, args: any(truthy(call_function(args[0], UNDEF, [float(x), float(i), t])) for i, x in enumerate(t.mv.tolist())),
    "every": lambda t, args: all(truthy(call_function(args[0], UNDEF, [float(x), float(i), t])) for i, x in enumerate(t.mv.tolist())),
    "reverse": lambda t, args: (t.mv.__setitem__(slice(None), array.array(t.code, t.mv.tolist()[::-1])), t)[1],
    "keys": lambda t, args: _PyIter(float(i) for i in range(t.length)),
    "values": lambda t, args: _PyIter(iterate(t)),
}


def _str_split(s, args):
    sep = _arg(args, 0)
    if sep is UNDEF:
        return JSArray([s])
    sep = to_str(sep)
    return JSArray(list(s) if sep == "" else s.split(sep))


STRING_METHODS = {
    "charAt": lambda s, args: s[int(to_num(_arg(args, 0, 0.0)))] if 0 <= int(to_num(_arg(args, 0, 0.0))) < len(s) else "",
    "charCodeAt": lambda s, args: float(ord(s[int(to_num(_arg(args, 0, 0.0)))])) if 0 <= int(to_num(_arg(args, 0, 0.0))) < len(s) else NAN,
    "indexOf": lambda s, args: float(s.find(to_str(_arg(args, 0)))),
    "includes": lambda s, args: to_str(_arg(args, 0)) in s,
    "startsWith": lambda s, args: s.startswith(to_str(_arg(args, 0))),
    "endsWith": lambda s, args: s.endswith(to_str(_arg(args, 0))),
    "slice": lambda s, args: s[_rel_index(_arg(args, 0), len(s), 0):_rel_index(_arg(args, 1), len(s), len(s))],
    "substring": lambda s, args: s[min(_rel_index(_arg(args, 0), len(s), 0), _rel_index(_arg(args, 1), len(s), len(s))):
                                   max(_rel_index(_arg(args, 0), len(s), 0), _rel_index(_arg(args, 1), len(s), len(s)))],
    "toUpperCase": lambda s, args: s.upper(),
    "toLowerCase": lambda s, args: s.lower(),
    "trim": lambda s, args: s.strip(),
    "split": _str_split,
    "padStart": lambda s, args: s.rjust(int(to_num(_arg(args, 0))), (to_str(_arg(args, 1)) if _arg(args, 1) is not UNDEF else " ")[:1] or " "),
    "padEnd": lambda s, args: s.ljust(int(to_num(_arg(args, 0))), (to_str(_arg(args, 1)) if _arg(args, 1) is not UNDEF else " ")[:1] or " "),
    "repeat": lambda s, args: s * int(to_num(_arg(args, 0))),
    "replace": lambda s, args: s.replace(to_str(_arg(args, 0)), to_str(_arg(args, 1)), 1),
    "toString": lambda s, args: s,
}


def _to_fixed(x, args):
    d = int(to_num(_arg(args, 0, 0.0)))
    if x != x:
        return "NaN"
    if abs(x) >= 1e21:
        return number_to_string(x)
    from decimal import ROUND_HALF_UP, Decimal
    q = Decimal(1).scaleb(-d)
    s = str(Decimal(x).quantize(q, rounding=ROUND_HALF_UP))
    if s.startswith("-") and float(s) == 0:
        s = s[1:] if x == 0 else s
    return s


def _to_exponential(x, args):
    d = _arg(args, 0)
    if x != x or x in (INF, -INF):
        return number_to_string(x)
    s = f"{x:.{int(to_num(d))}e}" if d is not UNDEF else repr(float(f"{x:.17e}"))
    if "e" not in s:
        s = f"{x:e}"
    mant, e = s.split("e")
    if d is UNDEF:
        mant = mant.rstrip("0").rstrip(".") if "." in mant else mant
    return f"{mant}e{'+' if int(e) >= 0 else '-'}{abs(int(e))}"


NUMBER_METHODS = {
    "toFixed": _to_fixed,
    "toExponential": _to_exponential,
    "toString": lambda x, args: number_to_string(x),
    "toPrecision": lambda x, args: f"{x:.{int(to_num(_arg(args, 0)))}g}",
}

SET_METHODS = {
    "add": lambda s, args: (s.add(_arg(args, 0)), s)[1],
    "has": lambda s, args: s.has(_arg(args, 0)),
    "delete": lambda s, args: s.delete(_arg(args, 0)),
    "clear": lambda s, args: (s.clear(), UNDEF)[1],
    "forEach": lambda s, args: ([call_function(args[0], UNDEF, [v, v, s]) for v in s.iterate()], UNDEF)[1],
    "values": lambda s, args: _PyIter(s.iterate()),
    "keys": lambda s, args: _PyIter(s.iterate()),
    "entries": lambda s, args: _PyIter(JSArray([v, v]) for v in s.iterate()),
}
MAP_METHODS = {
    "set": lambda m, args: (m.set(_arg(args, 0), _arg(args, 1)), m)[1],
    "get": lambda m, args: m.get(_arg(args, 0)),
    "has": lambda m, args: m.has(_arg(args, 0)),
    "delete": lambda m, args: m.delete(_arg(args, 0)),
    "clear": lambda m, args: (m.clear(), UNDEF)[1],
    "forEach": lambda m, args: ([call_function(args[0], UNDEF, [v, k, m]) for k, v in m.iterate()], UNDEF)[1],
    "keys": lambda m, args: _PyIter(k for k, v in m.iterate()),
    "values": lambda m, args: _PyIter(v for k, v in m.iterate()),
    "entries": lambda m, args: _PyIter(JSArray([k, v]) for k, v in m.iterate()),
}


def _math1(fn, domain_nan=True):
    def f(this, args):
        x = to_num(_arg(args, 0))
        try:
            return fn(x)
        except ValueError:
            return NAN
        except OverflowError:
            return INF
    return f


def _math_round(this, args):
    """round half toward +Infinity; floor(x + 0.5) would be wrong for 0.49999999999999994 and above 2^52"""
    x = to_num(_arg(args, 0))
    if x != x or x in (INF, -INF) or abs(x) >= 4503599627370496.0:
        return x
    f = math.floor(x)
    r = float(f + 1 if x - f >= 0.5 else f)          # x - floor(x) is exact below 2^52
    if r == 0 and (x < 0 or math.copysign(1.0, x) < 0):
        return -0.0
    return r


def _math_floor(this, args):
    x = to_num(_arg(args, 0))
    if x != x or x in (INF, -INF):
        return x
    r = float(math.floor(x))
    return -0.0 if r == 0 and math.copysign(1.0, x) < 0 else r


def _math_ceil(this, args):
    x = to_num(_arg(args, 0))
    if x != x or x in (INF, -INF):
        return x
    r = float(math.ceil(x))
    return -0.0 if r == 0 and (x < 0 or math.copysign(1.0, x) < 0) else r


def _math_trunc(this, args):
    x = to_num(_arg(args, 0))
    if x != x or x in (INF, -INF):
        return x
    r = float(math.trunc(x))
    return -0.0 if r == 0 and (x < 0 or math.copysign(1.0, x) < 0) else r


def _math_max(this, args):
    r = -INF
    for a in args:
        x = to_num(a)
        if x != x:
            return NAN
        if x > r or (x == 0 and r == 0 and math.copysign(1.0, r) < 0):
            r = x
    return r


def _math_min(this, args):
    r = INF
    for a in args:
        x = to_num(a)
        if x != x:
            return NAN
        if x < r or (x == 0 and r == 0 and math.copysign(1.0, x) < 0):
            r = x
    return r


def _math_log(x):
    if x == 0:
        return -INF
    return math.log(x)


def _math_sign(x):
    if x != x or x == 0:
        return x
    return 1.0 if x > 0 else -1.0


def _math_hypot(this, args):
    xs = [to_num(a) for a in args]
    if any(x in (INF, -INF) for x in xs):
        return INF
    if any(x != x for x in xs):
        return NAN
    return math.hypot(*xs)


def _math_atan2(this, args):
    return math.atan2(to_num(_arg(args, 0)), to_num(_arg(args, 1)))


def _math_imul(this, args):
    r = (to_int32(_arg(args, 0)) * to_int32(_arg(args, 1))) & 0xFFFFFFFF
    return float(r - 0x100000000 if r >= 0x80000000 else r)


def _fround(x):
    a = array.array("f", [0.0])
    try:
        a[0] = x
    except OverflowError:
        return INF if x > 0 else -INF
    return float(a[0])


def _math_cbrt(x):
    if x == 0 or x != x or x in (INF, -INF):
        return x
    return math.copysign(abs(x) ** (1.0 / 3.0), x) if not hasattr(math, "cbrt") else math.cbrt(x)


def _math_sqrt(x):
    if x < 0:
        return NAN
    return math.sqrt(x)


def make_math(random_fn=None):
    def no_random(this, args):
        throw_error("Error", "Math.random() is not available: every run must be seeded")
    m = {
        "PI": math.pi, "E": math.e, "LN2": math.log(2.0), "LN10": math.log(10.0), "LOG2E": 1 / math.log(2.0), "LOG10E": 1 / math.log(10.0),
        "SQRT2": math.sqrt(2.0), "SQRT1_2": math.sqrt(0.5),
        "floor": _math_floor, "ceil": _math_ceil, "round": _math_round, "trunc": _math_trunc, "max": _math_max, "min": _math_min,
        "abs": _math1(abs), "sqrt": _math1(_math_sqrt), "cbrt": _math1(_math_cbrt), "exp": _math1(math.exp), "log": _math1(_math_log),
        "log2": _math1(lambda x: -INF if x == 0 else math.log2(x)), "log10": _math1(lambda x: -INF if x == 0 else math.log10(x)),
        "log1p": _math1(math.log1p), "expm1": _math1(math.expm1),
        "sin": _math1(math.sin), "cos": _math1(math.cos), "tan": _math1(math.tan), "asin": _math1(math.asin), "acos": _math1(math.acos),
        "atan": _math1(math.atan), "sinh": _math1(math.sinh), "cosh": _math1(math.cosh), "tanh": _math1(math.tanh),
        "atan2": _math_atan2, "hypot": _math_hypot, "imul": _math_imul, "sign": _math1(_math_sign), "fround": _math1(_fround),
        "pow": lambda this, args: js_pow(to_num(_arg(args, 0)), to_num(_arg(args, 1))),
        "random": (lambda this, args: random_fn()) if random_fn else no_random,
    }
    obj = JSObject()
    for k, v in m.items():
        obj.props[k] = v if type(v) is float else HostFunction(v, k)
    return obj


def _typed_ctor(kind):
    def build(args):
        a = _arg(args, 0)
        if a is UNDEF:
            return JSTypedArray(kind, 0)
        if type(a) is float:
            if a < 0 or a != math.floor(a):
                throw_error("RangeError", f"Invalid typed array length: {to_str(a)}")
            return JSTypedArray(kind, int(a))
        if type(a) is JSTypedArray:
            if a.code == _TYPECODES[kind]:
                base = array.array(a.code, a.mv.tobytes())
                return JSTypedArray(kind, base=base, mv=memoryview(base))
            return _typed_from_iter(kind, (float(x) for x in a.mv.tolist()))
        if type(a) is ArrayBuffer:
            off = int(to_num(_arg(args, 1, 0.0)))
            code = _TYPECODES[kind]
            mv = memoryview(a.arr).cast("B")[off:]
            item = array.array(code).itemsize
            n = int(to_num(args[2])) if len(args) > 2 and args[2] is not UNDEF else len(mv) // item
            return JSTypedArray(kind, base=a.arr, mv=mv[:n * item].cast(code))
        if type(a) is JSObject and "length" in a.props:
            n = int(to_num(a.props["length"]))
            return _typed_from_iter(kind, (a.props.get(str(i), UNDEF) for i in range(n)))
        return _typed_from_iter(kind, iterate(a))

    def from_(this, args):
        src, fn = _arg(args, 0), _arg(args, 1)
        vals = list(iterate(src)) if not (type(src) is JSObject) else [UNDEF] * int(to_num(src.props.get("length", 0.0)))
        if fn is not UNDEF:
            vals = [call_function(fn, UNDEF, [v, float(i)]) for i, v in enumerate(vals)]
        return _typed_from_iter(kind, vals)
    f = HostFunction(lambda this, args: throw_error("TypeError", f"Constructor {kind} requires 'new'"), kind, build)
    f.props["from"] = HostFunction(from_, "from")
    f.props["BYTES_PER_ELEMENT"] = float(array.array(_TYPECODES[kind]).itemsize)
    return f


def _array_from(this, args):
    src, fn = _arg(args, 0), _arg(args, 1)
    if type(src) is JSObject:
        n = int(to_num(src.props.get("length", 0.0)))
        vals = [src.props.get(str(i), UNDEF) for i in range(n)]
    else:
        vals = list(iterate(src))
    if fn is not UNDEF:
        vals = [call_function(fn, UNDEF, [v, float(i)]) for i, v in enumerate(vals)]
    return JSArray(vals)


def _array_ctor(args):
    if len(args) == 1 and type(args[0]) is float:
        return JSArray([UNDEF] * int(args[0]))
    return JSArray(args)


def _object_assign(this, args):
    tgt = args[0]
    for src in args[1:]:
        if type(src) is JSObject:
            for k in own_keys(src):
                set_prop(tgt, k, get_prop(src, k))
    return tgt


def _entries_of(o):
    if type(o) is JSProxy:
        f = o.trap("ownKeys")
        keys = [to_str(k) for k in iterate(call_function(f, o.handler, [o.target]))] if f else [k for k, _ in _entries_of(o.target)]
        return [(k, get_prop(o, k)) for k in keys]
    if type(o) is JSObject:
        return [(k, get_prop(o, k)) for k in own_keys(o)]
    if type(o) is JSArray:
        return [(str(i), v) for i, v in enumerate(o)]
    if type(o) is JSTypedArray:
        return [(str(i), float(v)) for i, v in enumerate(o.mv.tolist())]
    return []


def _parse_float(this, args):
    s = to_str(_arg(args, 0)).strip()
    import re
    m = re.match(r"[+-]?(Infinity|(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?))", s)
    return float(m.group(0).replace("Infinity", "inf")) if m else NAN


def _parse_int(this, args):
    s = to_str(_arg(args, 0)).strip()
    radix = int(to_num(_arg(args, 1, 10.0))) or 10
    import re
    if radix == 16 or (len(args) < 2 and s[:2].lower() == "0x"):
        m = re.match(r"[+-]?(0[xX])?[0-9a-fA-F]+", s)
        return float(int(m.group(0), 16)) if m else NAN
    digits = "0123456789abcdefghijklmnopqrstuvwxyz"[:radix]
    m = re.match(r"[+-]?[" + digits + digits.upper() + r"]+", s)
    return float(int(m.group(0), radix)) if m else NAN


def _json_value(v, indent=None):
    t = type(v)
    if t is float:
        return number_to_string(v) if v == v and v not in (INF, -INF) else "null"
    if t is str:
        import json
        return json.dumps(v)
    if t is bool:
        return "true" if v else "false"
    if v is None:
        return "null"
    if t is JSArray:
        return "[" + ",".join(_json_value(x) if x is not UNDEF else "null" for x in v) + "]"
    if t is JSTypedArray:
        return "{" + ",".join(f'"{i}":{_json_value(float(x))}' for i, x in enumerate(v.mv.tolist())) + "}"
    if t is JSObject:
        import json
        return "{" + ",".join(f"{json.dumps(k)}:{_json_value(get_prop(v, k))}" for k in own_keys(v)
                              if type(get_prop(v, k)) not in (JSFunction, HostFunction, _Undefined)) + "}"
    return "null"


def make_globals(log=None, random_fn=None):
    g = {}
    g["undefined"] = UNDEF
    g["NaN"] = NAN
    g["Infinity"] = INF
    g["Math"] = make_math(random_fn)
    for kind in _TYPECODES:
        g[kind] = _typed_ctor(kind)
    arr = HostFunction(lambda this, args: _array_ctor(args), "Array", _array_ctor)
    arr.props["from"] = HostFunction(_array_from, "from")
    arr.props["isArray"] = HostFunction(lambda this, args: type(_arg(args, 0)) is JSArray, "isArray")
    arr.props["of"] = HostFunction(lambda this, args: JSArray(args), "of")
    g["Array"] = arr

    def set_ctor(args):
        s = JSSet()
        a = _arg(args, 0)
        if a is not UNDEF and a is not None:
            for v in iterate(a):
                s.add(v)
        return s

    def map_ctor(args):
        m = JSMap()
        a = _arg(args, 0)
        if a is not UNDEF and a is not None:
            for kv in iterate(a):
                m.set(get_prop(kv, 0.0), get_prop(kv, 1.0))
        return m
    g["Set"] = HostFunction(lambda this, args: throw_error("TypeError", "Constructor Set requires 'new'"), "Set", set_ctor)
    g["Map"] = HostFunction(lambda this, args: throw_error("TypeError", "Constructor Map requires 'new'"), "Map", map_ctor)
    obj = HostFunction(lambda this, args: JSObject(), "Object", lambda args: JSObject())
    obj.props["keys"] = HostFunction(lambda this, args: JSArray(k for k, v in _entries_of(args[0])), "keys")
    obj.props["values"] = HostFunction(lambda this, args: JSArray(v for k, v in _entries_of(args[0])), "values")
    obj.props["entries"] = HostFunction(lambda this, args: JSArray(JSArray([k, v]) for k, v in _entries_of(args[0])), "entries")
    obj.props["assign"] = HostFunction(_object_assign, "assign")
    obj.props["freeze"] = HostFunction(lambda this, args: args[0], "freeze")
    obj.props["fromEntries"] = HostFunction(lambda this, args: JSObject(None, {prop_key(get_prop(kv, 0.0)): get_prop(kv, 1.0) for kv in iterate(args[0])}), "fromEntries")
    g["Object"] = obj
    num = HostFunction(lambda this, args: to_num(_arg(args, 0, 0.0)), "Number")
    num.props.update({"isFinite": HostFunction(lambda this, args: type(_arg(args, 0)) is float and args[0] == args[0] and args[0] not in (INF, -INF), "isFinite"),
                      "isNaN": HostFunction(lambda this, args: type(_arg(args, 0)) is float and args[0] != args[0], "isNaN"),
                      "isInteger": HostFunction(lambda this, args: type(_arg(args, 0)) is float and args[0] == args[0] and args[0] not in (INF, -INF) and args[0] == math.floor(args[0]), "isInteger"),
                      "parseFloat": HostFunction(_parse_float, "parseFloat"), "parseInt": HostFunction(_parse_int, "parseInt"),
                      "MAX_SAFE_INTEGER": 9007199254740991.0, "MIN_SAFE_INTEGER": -9007199254740991.0, "EPSILON": 2.0 ** -52,
                      "MAX_VALUE": 1.7976931348623157e308, "MIN_VALUE": 5e-324, "POSITIVE_INFINITY": INF, "NEGATIVE_INFINITY": -INF, "NaN": NAN})
    g["Number"] = num
    g["String"] = HostFunction(lambda this, args: to_str(_arg(args, 0, "")), "String")
    g["Boolean"] = HostFunction(lambda this, args: truthy(_arg(args, 0)), "Boolean")
    g["isNaN"] = HostFunction(lambda this, args: to_num(_arg(args, 0)) != to_num(_arg(args, 0)), "isNaN")
    g["isFinite"] = HostFunction(lambda this, args: (lambda x: x == x and x not in (INF, -INF))(to_num(_arg(args, 0))), "isFinite")
    g["parseFloat"] = HostFunction(_parse_float, "parseFloat")
    g["parseInt"] = HostFunction(_parse_int, "parseInt")

    def error_ctor(kind):
        def build(args):
            msg = to_str(_arg(args, 0)) if _arg(args, 0) is not UNDEF else ""
            return make_error(kind, msg)
        return HostFunction(lambda this, args: build(args), kind, build)
    for kind in ("Error", "TypeError", "RangeError"):
        g[kind] = error_ctor(kind)
    lines = log if log is not None else []

    def console(level):
        return HostFunction(lambda this, args: (lines.append(level + ": " + " ".join(to_display(a) for a in args)), UNDEF)[1], level)
    g["console"] = JSObject(None, {k: console(k) for k in ("log", "warn", "error", "info", "debug", "time", "timeEnd")})
    t0 = time.perf_counter()
    g["performance"] = JSObject(None, {"now": HostFunction(lambda this, args: (time.perf_counter() - t0) * 1000.0, "now")})
    g["Date"] = JSObject(None, {"now": HostFunction(lambda this, args: float(int(time.time() * 1000)), "now")})
    g["JSON"] = JSObject(None, {"stringify": HostFunction(lambda this, args: _json_value(_arg(args, 0)), "stringify")})
    g["globalThis"] = JSObject()

    def proxy_ctor(args):
        if type(_arg(args, 0)) is not JSObject or type(_arg(args, 1)) is not JSObject:
            throw_error("TypeError", "Cannot create proxy with a non-object as target or handler")
        return JSProxy(args[0], args[1])
    g["Proxy"] = HostFunction(lambda this, args: throw_error("TypeError", "Constructor Proxy requires 'new'"), "Proxy", proxy_ctor)
    return g


# ---------------------------------------------------------------------------------------------------------------------
# compiler: AST → closures over Env
# ---------------------------------------------------------------------------------------------------------------------


class CScope:
    __slots__ = ("names", "parent", "is_function")

    def __init__(self, parent, is_function):
        self.names, self.parent, self.is_function = set(), parent, is_function


def pattern_names(p, out):
    k = p[0]
    if k == "id":
        out.append(p[1])
    elif k == "assignpat":
        pattern_names(p[1], out)
    elif k == "objpat":
        for _, v in p[1]:
            pattern_names(v, out)
        if p[2]:
            out.append(p[2])
    elif k == "arrpat":
        for e in p[1]:
            if e is not None:
                pattern_names(e, out)
        if p[2]:
            pattern_names(p[2], out)
    return out


def hoisted_vars(stmts, out):
    """names declared with `var` anywhere in these statements (not inside nested functions)"""
    for s in stmts:
        if s is None:
            continue
        k = s[0]
        if k == "var" and s[1] == "var":
            for target, _ in s[2]:
                pattern_names(target, out)
        elif k == "block":
            hoisted_vars(s[1], out)
        elif k == "if":
            hoisted_vars([s[2], s[3]], out)
        elif k == "for":
            hoisted_vars([s[1], s[4]], out)
        elif k in ("forof", "forin"):
            if s[1] == "var":
                pattern_names(s[2], out)
            hoisted_vars([s[4]], out)
        elif k == "while":
            hoisted_vars([s[2]], out)
        elif k == "dowhile":
            hoisted_vars([s[1]], out)
        elif k == "try":
            hoisted_vars([s[1], s[3], s[4]], out)
        elif k == "switch":
            for _, body in s[2]:
                hoisted_vars(body, out)
        elif k == "export" and s[1] == "decl":
            hoisted_vars([s[2]], out)
    return out


def lexical_names(stmts):
    """let / const / class / function declarations made directly by these statements"""
    out = []
    for s in stmts:
        k = s[0]
        if k == "export" and s[1] in ("decl", "default"):
            s = s[2]
            k = s[0]
        if k == "var" and s[1] != "var":
            for target, _ in s[2]:
                pattern_names(target, out)
        elif k in ("funcdecl", "classdecl"):
            if s[1]:
                out.append(s[1])
    return out


def contains_fn(node) -> bool:
    if isinstance(node, tuple):
        if node and node[0] in ("fn", "class"):
            return True
        return any(contains_fn(x) for x in node)
    if isinstance(node, list):
        return any(contains_fn(x) for x in node)
    return False


class Compiler:
    def __init__(self, interp, fname):
        self.interp = interp
        self.fname = fname
        self.globals = interp.globals

    # ---- scope resolution ----
    def resolve(self, name, scope):
        hops = 0
        s = scope
        while s is not None:
            if name in s.names:
                return hops
            s = s.parent
            hops += 1
        return -1

    def getter(self, name, scope, line=None):
        hops = self.resolve(name, scope)
        if hops == 0:
            return lambda env: env.v[name]
        if hops == 1:
            return lambda env: env.p.v[name]
        if hops == 2:
            return lambda env: env.p.p.v[name]
        if hops == 3:
            return lambda env: env.p.p.p.v[name]
        if hops > 3:
            def get(env):
                for _ in range(hops):
                    env = env.p
                return env.v[name]
            return get
        g = self.globals
        fname = self.fname

        def get_global(env):
            try:
                return g[name]
            except KeyError:
                throw_error("ReferenceError", f"{name} is not defined ({os.path.basename(fname)})")
        return get_global

    def setter(self, name, scope):
        hops = self.resolve(name, scope)
        if hops < 0:
            g = self.globals

            def set_global(env, value):
                g[name] = value
            return set_global
        if hops == 0:
            def set0(env, value):
                env.v[name] = value
            return set0
        if hops == 1:
            def set1(env, value):
                env.p.v[name] = value
            return set1

        def setn(env, value):
            for _ in range(hops):
                env = env.p
            env.v[name] = value
        return setn

    # ---- patterns ----
    def binder(self, pattern, scope, declare):
        """closure(env, value) that binds / assigns the pattern; declare: names live in `scope` itself (hops 0)"""
        k = pattern[0]
        if k == "id":
            name = pattern[1]
            if declare:
                def bind_id(env, value):
                    env.v[name] = value
                return bind_id
            return self.setter(name, scope)
        if k == "member":
            objf = self.expr(pattern[1], scope)
            keyf = self.expr(pattern[2], scope)
            return lambda env, value: set_prop(objf(env), keyf(env), value)
        if k == "assignpat":
            inner = self.binder(pattern[1], scope, declare)
            dflt = self.expr(pattern[2], scope)

            def bind_default(env, value):
                inner(env, dflt(env) if value is UNDEF else value)
            return bind_default
        if k == "objpat":
            parts = [(key, self.binder(p, scope, declare)) for key, p in pattern[1]]
            rest = pattern[2]
            rest_set = None
            if rest:
                rest_set = self.binder(("id", rest), scope, declare)
            taken = [key for key, _ in pattern[1]]

            def bind_obj(env, value):
                if value is None or value is UNDEF:
                    throw_error("TypeError", f"Cannot destructure {to_display(value)}")
                for key, b in parts:
                    b(env, get_prop(value, key))
                if rest_set:
                    o = JSObject()
                    for kk, vv in _entries_of(value):
                        if kk not in taken:
                            o.props[kk] = vv
                    rest_set(env, o)
            return bind_obj
        if k == "arrpat":
            parts = [None if p is None else self.binder(p, scope, declare) for p in pattern[1]]
            rest = self.binder(pattern[2], scope, declare) if pattern[2] else None

            def bind_arr(env, value):
                vals = list(iterate(value)) if rest or type(value) not in (JSArray, JSTypedArray) else None
                if vals is None:
                    for i, b in enumerate(parts):
                        if b:
                            b(env, get_prop(value, float(i)))
                    return
                for i, b in enumerate(parts):
                    if b:
                        b(env, vals[i] if i < len(vals) else UNDEF)
                if rest:
                    rest(env, JSArray(vals[len(parts):]))
            return bind_arr
        raise NotImplementedError(f"pattern {k}")

    # ---- functions ----
    def function(self, node, scope):
        _, name, params, rest, body, is_arrow, expr_body, line = node
        fscope = CScope(scope, True)
        pnames = []
        for p in params:
            pattern_names(p, pnames)
        if rest:
            pattern_names(rest, pnames)
        fscope.names.update(pnames)
        fscope.names.add("this") if not is_arrow else None
        var_names = []
        if not expr_body:
            hoisted_vars(body[1], var_names)
            fscope.names.update(var_names)
            fscope.names.update(lexical_names(body[1]))
        if not is_arrow:
            fscope.names.add("arguments")
        binders = [self.binder(p, fscope, True) for p in params]
        simple = all(p[0] == "id" for p in params)
        simple_names = [p[1] for p in params] if simple else None
        rest_b = self.binder(rest, fscope, True) if rest else None
        n = len(params)

        if simple and not rest:
            def bind_params(env, args):
                v = env.v
                la = len(args)
                for i in range(n):
                    v[simple_names[i]] = args[i] if i < la else UNDEF
        else:
            def bind_params(env, args):
                la = len(args)
                for i in range(n):
                    binders[i](env, args[i] if i < la else UNDEF)
                if rest_b:
                    rest_b(env, JSArray(args[n:]))
        uses_arguments = (not is_arrow) and self._uses_identifier(body, "arguments")
        if uses_arguments:
            inner_bind = bind_params

            def bind_params(env, args):       # noqa: F811
                env.v["arguments"] = JSArray(args)
                inner_bind(env, args)
        if expr_body:
            bodyf = self.expr(body, fscope)
        else:
            bodyf = self.statements(body[1], fscope, function_level=True)
        fname = self.fname
        var_names = tuple(dict.fromkeys(var_names))
        this_get = None
        if is_arrow:
            this_get = self.getter("this", scope) if self.resolve("this", scope) >= 0 else (lambda env: UNDEF)

        def make(env):
            f = JSFunction()
            f.name = name or ""
            f.nparams = n
            f.bind_params = bind_params
            f.body = bodyf
            f.env = env
            f.is_arrow = is_arrow
            f.expr_body = expr_body
            f.line = line
            f.fname = fname
            f.var_names = var_names
            if is_arrow:
                f.this_val = this_get(env)
            else:
                f.props["prototype"] = JSObject(None, {"constructor": f})
            return f
        return make

    def _uses_identifier(self, node, name):
        if isinstance(node, tuple):
            if len(node) == 2 and node[0] == "id" and node[1] == name:
                return True
            if node and node[0] == "fn" and not node[5]:
                return False
            return any(self._uses_identifier(x, name) for x in node)
        if isinstance(node, list):
            return any(self._uses_identifier(x, name) for x in node)
        return False

    def class_(self, node, scope):
        _, name, sup, members = node
        if sup is not None:
            raise NotImplementedError("class inheritance is not supported")
        ctor_node = None
        methods, fields, statics = [], [], []
        for kind, static, key, fn in members:
            if kind == "method" and key == "constructor" and not static:
                ctor_node = fn
            elif kind == "field":
                (statics if static else fields).append((key, self.expr(fn, CScope(scope, True)) if fn else None, "field"))
            else:
                (statics if static else methods).append((key, self.function(fn, scope), kind))
        if ctor_node is None:
            ctor_node = ("fn", name, [], None, ("block", []), False, False, 0)
        make_ctor = self.function(("fn", name) + ctor_node[2:], scope)

        def make(env):
            f = make_ctor(env)
            f.is_class = True
            proto = f.props["prototype"]
            for key, mk, kind in methods:
                fn = mk(env)
                if kind == "method":
                    proto.props[key] = fn
                else:
                    acc = proto.props.get(key)
                    if type(acc) is not Accessor:
                        acc = proto.props[key] = Accessor()
                    setattr(acc, kind, fn)
            f.fields = [(key, init) for key, init, _ in fields] or None
            for key, mk, kind in statics:
                f.props[key] = mk(Env({"this": f}, env)) if kind == "field" and mk else (mk(env) if mk else UNDEF)
            return f
        return make

    # ---- statements ----
    def statements(self, stmts, scope, function_level=False):
        """compiled statement list; declares its lexical names in `scope` when function_level, else in a new block scope if any"""
        lex = lexical_names(stmts)
        new_scope = bool(lex) and not function_level
        inner = CScope(scope, False) if new_scope else scope
        if new_scope:
            inner.names.update(lex)
        # function declarations are initialised at block entry
        fdecls = []
        body = []
        for s in stmts:
            t = s
            if t[0] == "export" and t[1] in ("decl", "default") and t[2][0] == "funcdecl":
                t = t[2]
            if t[0] == "funcdecl":
                fdecls.append((t[1], self.function(t[2], inner)))
        for s in stmts:
            c = self.statement(s, inner)
            if c is not None:
                body.append(c)
        body = tuple(body)
        fdecls = tuple(fdecls)

        if len(body) == 1 and not fdecls and not new_scope:
            return body[0]

        def run(env):
            if new_scope:
                env = Env({}, env)
            for fname_, mk in fdecls:
                env.v[fname_] = mk(env)
            for st in body:
                r = st(env)
                if r is not None:
                    return r
            return None
        return run

    def statement(self, s, scope):
        k = s[0]
        if k == "expr":
            e = self.expr(s[1], scope)

            def run_expr(env):
                e(env)
            return run_expr
        if k == "var":
            kind = s[1]
            parts = []
            for target, init in s[2]:
                initf = self.expr(init, scope) if init is not None else None
                if kind == "var":
                    b = self.binder(target, scope, False)
                    if initf is None:
                        continue
                else:
                    b = self.binder(target, scope, True)
                if target[0] == "id" and kind != "var":
                    parts.append((target[1], initf))
                else:
                    parts.append((b, initf))
            if len(parts) == 1 and type(parts[0][0]) is str:
                name, initf = parts[0]
                if initf is None:
                    def decl1u(env):
                        env.v[name] = UNDEF
                    return decl1u

                def decl1(env):
                    env.v[name] = initf(env)
                return decl1

            def decl(env):
                for b, initf in parts:
                    v = initf(env) if initf else UNDEF
                    if type(b) is str:
                        env.v[b] = v
                    else:
                        b(env, v)
            return decl
        if k == "funcdecl":
            return None        # hoisted by statements()
        if k == "classdecl":
            mk = self.class_(s[2], scope)
            name = s[1]

            def declc(env):
                env.v[name] = mk(env)
            return declc
        if k == "return":
            if s[1] is None:
                r0 = Ret(UNDEF)
                return lambda env: r0
            e = self.expr(s[1], scope)
            return lambda env: Ret(e(env))
        if k == "if":
            test = self.expr(s[1], scope)
            cons = self.statements([s[2]], scope) if s[2][0] != "block" else self.statements(s[2][1], scope)
            alt = None
            if s[3] is not None:
                alt = self.statements([s[3]], scope) if s[3][0] != "block" else self.statements(s[3][1], scope)
            if alt is None:
                def if1(env):
                    v = test(env)
                    if v is True or (v is not False and truthy(v)):
                        return cons(env)
                return if1

            def if2(env):
                v = test(env)
                if v is True or (v is not False and truthy(v)):
                    return cons(env)
                return alt(env)
            return if2
        if k == "block":
            return self.statements(s[1], scope)
        if k == "for":
            return self.for_(s, scope)
        if k == "forof" or k == "forin":
            return self.for_of(s, scope)
        if k == "while":
            test = self.expr(s[1], scope)
            body = self.loop_body(s[2], scope)

            def while_(env):
                while True:
                    v = test(env)
                    if not (v is True or (v is not False and truthy(v))):
                        return None
                    r = body(env)
                    if r is not None:
                        if r is BRK:
                            return None
                        if r is not CNT:
                            return r
            return while_
        if k == "dowhile":
            test = self.expr(s[2], scope)
            body = self.loop_body(s[1], scope)

            def dowhile(env):
                while True:
                    r = body(env)
                    if r is not None:
                        if r is BRK:
                            return None
                        if r is not CNT:
                            return r
                    if not truthy(test(env)):
                        return None
            return dowhile
        if k == "break":
            return lambda env: BRK
        if k == "continue":
            return lambda env: CNT
        if k == "empty":
            return None
        if k == "throw":
            e = self.expr(s[1], scope)

            def throw(env):
                raise JSThrow(e(env))
            return throw
        if k == "try":
            block = self.statements(s[1][1], scope)
            handler = None
            if s[3] is not None:
                hscope = CScope(scope, False)
                names = pattern_names(s[2], []) if s[2] else []
                hscope.names.update(names)
                bind = self.binder(s[2], hscope, True) if s[2] else None
                hbody = self.statements(s[3][1], hscope)
            final = self.statements(s[4][1], scope) if s[4] is not None else None
            has_handler = s[3] is not None

            def try_(env):
                try:
                    try:
                        return block(env)
                    except JSThrow as ex:
                        if not has_handler:
                            raise
                        henv = Env({}, env)
                        if bind:
                            bind(henv, ex.value)
                        return hbody(henv)
                finally:
                    if final is not None:
                        r = final(env)
                        if r is not None:
                            return r      # noqa: B012 — JS semantics: a completion in finally wins
            return try_
        if k == "switch":
            disc = self.expr(s[1], scope)
            sscope = CScope(scope, False)
            allstmts = [st for _, body in s[2] for st in body]
            sscope.names.update(lexical_names(allstmts))
            cases = [(self.expr(t, sscope) if t is not None else None, tuple(c for c in (self.statement(st, sscope) for st in body) if c is not None))
                     for t, body in s[2]]

            def switch(env):
                v = disc(env)
                env = Env({}, env)
                start = None
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_equals(t(env), v):
                        start = i
                        break
                if start is None:
                    for i, (t, _) in enumerate(cases):
                        if t is None:
                            start = i
                            break
                if start is None:
                    return None
                for _, body in cases[start:]:
                    for st in body:
                        r = st(env)
                        if r is not None:
                            if r is BRK:
                                return None
                            return r
                return None
            return switch
        if k == "import":
            return None        # handled by the module loader
        if k == "export":
            if s[1] == "decl":
                return self.statement(s[2], scope)
            if s[1] == "default":
                return self.statement(s[2], scope)
            if s[1] == "defaultexpr":
                e = self.expr(s[2], scope)

                def set_default(env):
                    env.v["*default*"] = e(env)
                return set_default
            return None
        raise NotImplementedError(f"statement {k}")

    def loop_body(self, body, scope):
        if body[0] == "block":
            return self.statements(body[1], scope)
        return self.statements([body], scope)

    def for_(self, s, scope):
        _, init, test, update, body = s
        lscope = scope
        names = []
        if init is not None and init[0] == "var" and init[1] != "var":
            for target, _ in init[2]:
                pattern_names(target, names)
            lscope = CScope(scope, False)
            lscope.names.update(names)
        initf = self.statement(init, lscope) if init is not None else None
        testf = self.expr(test, lscope) if test is not None else None
        updatef = self.expr(update, lscope) if update is not None else None
        bodyf = self.loop_body(body, lscope)
        new_env = lscope is not scope
        per_iter = new_env and contains_fn(body)

        def for_run(env):
            if new_env:
                env = Env({}, env)
            if initf:
                initf(env)
            while True:
                if testf is not None:
                    v = testf(env)
                    if not (v is True or (v is not False and truthy(v))):
                        return None
                r = bodyf(env)
                if r is not None:
                    if r is BRK:
                        return None
                    if r is not CNT:
                        return r
                if per_iter:
                    env = Env(dict(env.v), env.p)
                if updatef:
                    updatef(env)
        return for_run

    def for_of(self, s, scope):
        k, kind, target, it, body = s
        itf = self.expr(it, scope)
        if kind in ("let", "const"):
            lscope = CScope(scope, False)
            lscope.names.update(pattern_names(target, []))
            bind = self.binder(target, lscope, True)
            new_env = True
        else:
            lscope = scope
            bind = self.binder(target, scope, False)
            new_env = False
        bodyf = self.loop_body(body, lscope)
        per_iter = new_env and contains_fn(body)
        simple = target[1] if (target[0] == "id" and new_env) else None
        is_in = k == "forin"

        def forof(env):
            src = itf(env)
            if is_in:
                if src is None or src is UNDEF:
                    return None
                seq = [kk for kk, _ in _entries_of(src)]
            else:
                seq = iterate(src)
            if new_env:
                env = Env({}, env)
            for v in seq:
                if per_iter:
                    env = Env({}, env.p)
                if simple is not None:
                    env.v[simple] = v
                else:
                    bind(env, v)
                r = bodyf(env)
                if r is not None:
                    if r is BRK:
                        return None
                    if r is not CNT:
                        return r
            return None
        return forof

    # ---- expressions ----
    def expr(self, e, scope):
        k = e[0]
        m = getattr(self, "e_" + k, None)
        if m is None:
            raise NotImplementedError(f"expression {k}")
        return m(e, scope)

    def e_num(self, e, scope):
        v = e[1]
        return lambda env: v

    def e_str(self, e, scope):
        v = e[1]
        return lambda env: v

    def e_bool(self, e, scope):
        v = e[1]
        return lambda env: v

    def e_null(self, e, scope):
        return lambda env: None

    def e_paren(self, e, scope):
        return self.expr(e[1], scope)

    def e_this(self, e, scope):
        if self.resolve("this", scope) < 0:
            return lambda env: UNDEF
        return self.getter("this", scope)

    def e_id(self, e, scope):
        return self.getter(e[1], scope)

    def e_tpl(self, e, scope):
        parts = [p if isinstance(p, str) else self.expr(p, scope) for p in e[1]]
        return lambda env: "".join(p if type(p) is str else to_str(p(env)) for p in parts)

    def e_arr(self, e, scope):
        items = [None if x is None else (("s", self.expr(x[1], scope)) if x[0] == "spread" else ("v", self.expr(x, scope))) for x in e[1]]
        if all(it is not None and it[0] == "v" for it in items):
            fs = [it[1] for it in items]
            return lambda env: JSArray([f(env) for f in fs])

        def arr(env):
            out = JSArray()
            for it in items:
                if it is None:
                    out.append(UNDEF)
                elif it[0] == "v":
                    out.append(it[1](env))
                else:
                    out.extend(iterate(it[1](env)))
            return out
        return arr

    def e_obj(self, e, scope):
        props = []
        for p in e[1]:
            if p[0] == "spread":
                props.append(("spread", None, self.expr(p[1], scope)))
            elif p[0] == "prop":
                key = p[1][1] if p[1][0] == "str" else self.expr(p[1], scope)
                props.append(("prop", key, self.expr(p[2], scope)))
            else:
                props.append((p[0], p[1][1], self.function(p[2], scope)))

        def obj(env):
            o = JSObject()
            d = o.props
            for kind, key, f in props:
                if kind == "prop":
                    d[key if type(key) is str else prop_key(key(env))] = f(env)
                elif kind == "spread":
                    src = f(env)
                    for kk, vv in _entries_of(src):
                        d[kk] = vv
                else:
                    acc = d.get(key)
                    if type(acc) is not Accessor:
                        acc = d[key] = Accessor()
                    setattr(acc, kind, f(env))
            return o
        return obj

    def e_fn(self, e, scope):
        return self.function(e, scope)

    def e_class(self, e, scope):
        return self.class_(e, scope)

    def e_seq(self, e, scope):
        fs = [self.expr(x, scope) for x in e[1]]

        def seq(env):
            v = UNDEF
            for f in fs:
                v = f(env)
            return v
        return seq

    def e_cond(self, e, scope):
        t, a, b = self.expr(e[1], scope), self.expr(e[2], scope), self.expr(e[3], scope)

        def cond(env):
            v = t(env)
            if v is True or (v is not False and truthy(v)):
                return a(env)
            return b(env)
        return cond

    def e_logical(self, e, scope):
        op = e[1]
        lf, rf = self.expr(e[2], scope), self.expr(e[3], scope)
        if op == "&&":
            def and_(env):
                v = lf(env)
                if v is True or (v is not False and truthy(v)):
                    return rf(env)
                return v
            return and_
        if op == "||":
            def or_(env):
                v = lf(env)
                if v is True or (v is not False and truthy(v)):
                    return v
                return rf(env)
            return or_

        def nullish(env):
            v = lf(env)
            return rf(env) if (v is None or v is UNDEF) else v
        return nullish

    def e_unary(self, e, scope):
        op = e[1]
        if op == "typeof":
            if e[2][0] == "id" and self.resolve(e[2][1], scope) < 0:
                name, g = e[2][1], self.globals
                return lambda env: typeof(g[name]) if name in g else "undefined"
            f = self.expr(e[2], scope)
            return lambda env: typeof(f(env))
        if op == "delete":
            if e[2][0] != "member":
                raise NotImplementedError("delete of a non-member")
            objf, keyf = self.expr(e[2][1], scope), self.expr(e[2][2], scope)

            def delete(env):
                o = objf(env)
                if type(o) is JSObject:
                    o.props.pop(prop_key(keyf(env)), None)
                return True
            return delete
        f = self.expr(e[2], scope)
        if op == "-":
            def neg(env):
                v = f(env)
                return -v if type(v) is float else -to_num(v)
            return neg
        if op == "+":
            return lambda env: to_num(f(env))
        if op == "!":
            def not_(env):
                v = f(env)
                if v is True:
                    return False
                if v is False:
                    return True
                return not truthy(v)
            return not_
        if op == "~":
            return lambda env: float(~to_int32(f(env)))
        if op == "void":
            return lambda env: (f(env), UNDEF)[1]
        raise NotImplementedError(op)

    def e_bin(self, e, scope):
        op = e[1]
        lf, rf = self.expr(e[2], scope), self.expr(e[3], scope)
        if op == "+":
            def add(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a + b
                return js_add(a, b)
            return add
        if op == "-":
            def sub(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a - b
                return to_num(a) - to_num(b)
            return sub
        if op == "*":
            def mul(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a * b
                return to_num(a) * to_num(b)
            return mul
        if op == "/":
            def div(env):
                a = lf(env); b = rf(env)
                if type(a) is not float:
                    a = to_num(a)
                if type(b) is not float:
                    b = to_num(b)
                if b != 0.0:
                    return a / b
                return js_div(a, b)
            return div
        if op == "%":
            def mod(env):
                a = lf(env); b = rf(env)
                if type(a) is not float:
                    a = to_num(a)
                if type(b) is not float:
                    b = to_num(b)
                if b != 0.0 and a == a and a not in (INF, -INF) and b == b:
                    if b in (INF, -INF):
                        return a
                    return math.fmod(a, b)
                return js_mod(a, b)
            return mod
        if op == "**":
            return lambda env: js_pow(to_num(lf(env)), to_num(rf(env)))
        if op == "<":
            def lt(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a < b
                return js_less(a, b)
            return lt
        if op == ">":
            def gt(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a > b
                return js_less(b, a)
            return gt
        if op == "<=":
            def le(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a <= b
                return js_less(a, b, orequal=True)
            return le
        if op == ">=":
            def ge(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a >= b
                return js_less(b, a, orequal=True)
            return ge
        if op == "===":
            def seq_(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a == b
                return strict_equals(a, b)
            return seq_
        if op == "!==":
            def sne(env):
                a = lf(env); b = rf(env)
                if type(a) is float and type(b) is float:
                    return a != b
                return not strict_equals(a, b)
            return sne
        if op == "==":
            return lambda env: loose_equals(lf(env), rf(env))
        if op == "!=":
            return lambda env: not loose_equals(lf(env), rf(env))
        if op == "|":
            return lambda env: float(_s32(to_int32(lf(env)) | to_int32(rf(env))))
        if op == "&":
            return lambda env: float(_s32(to_int32(lf(env)) & to_int32(rf(env))))
        if op == "^":
            return lambda env: float(_s32(to_int32(lf(env)) ^ to_int32(rf(env))))
        if op == "<<":
            return lambda env: float(_s32((to_int32(lf(env)) << (to_uint32(rf(env)) & 31)) & 0xFFFFFFFF))
        if op == ">>":
            return lambda env: float(to_int32(lf(env)) >> (to_uint32(rf(env)) & 31))
        if op == ">>>":
            return lambda env: float(to_uint32(lf(env)) >> (to_uint32(rf(env)) & 31))
        if op == "in":
            def in_(env):
                key, o = lf(env), rf(env)
                if type(o) is JSObject:
                    k2 = prop_key(key)
                    while o is not None:
                        if k2 in o.props:
                            return True
                        o = o.proto
                    return False
                if type(o) in (JSArray, JSTypedArray):
                    i = index_of(key)
                    n = len(o) if type(o) is JSArray else o.length
                    return 0 <= i < n or key == "length"
                if type(o) is JSProxy:
                    f = o.trap("has")
                    if f:
                        return truthy(call_function(f, o.handler, [o.target, prop_key(key)]))
                    return prop_key(key) in o.target.props
                throw_error("TypeError", "Cannot use 'in' operator on a primitive")
            return in_
        if op == "instanceof":
            def instanceof(env):
                o, c = lf(env), rf(env)
                if type(c) is HostFunction:
                    return {"Array": type(o) is JSArray, "Set": type(o) is JSSet, "Map": type(o) is JSMap,
                            "Error": type(o) is JSObject and o.cls == "Error", "Object": type(o) in (JSObject, JSArray)}.get(
                                c.name, type(o) is JSTypedArray and o.kind == c.name)
                if type(c) is JSFunction and type(o) is JSObject:
                    p = o.proto
                    target = c.props.get("prototype")
                    while p is not None:
                        if p is target:
                            return True
                        p = p.proto
                return False
            return instanceof
        raise NotImplementedError(f"operator {op}")

    BINOPS = {"+=": "+", "-=": "-", "*=": "*", "/=": "/", "%=": "%", "**=": "**", "<<=": "<<", ">>=": ">>", ">>>=": ">>>", "&=": "&", "|=": "|",
              "^=": "^"}

    def e_assign(self, e, scope):
        _, op, target, value = e
        vf = self.expr(value, scope)
        if op == "=":
            if target[0] == "id":
                setf = self.setter(target[1], scope)

                def assign_id(env):
                    v = vf(env)
                    setf(env, v)
                    return v
                return assign_id
            if target[0] == "member":
                objf, keyf = self.expr(target[1], scope), self.expr(target[2], scope)

                def assign_member(env):
                    o = objf(env)
                    k = keyf(env)
                    v = vf(env)
                    if type(o) is JSTypedArray and type(k) is float:
                        i = int(k)
                        if i == k:
                            o.store(i, v)
                        return v
                    set_prop(o, k, v)
                    return v
                return assign_member
            bind = self.binder(target, scope, False)

            def assign_pattern(env):
                v = vf(env)
                bind(env, v)
                return v
            return assign_pattern
        if op in ("&&=", "||=", "??="):
            getf = self.expr(target, scope)
            setf = self.binder(target, scope, False)

            def assign_logical(env):
                cur = getf(env)
                if op == "&&=" and not truthy(cur):
                    return cur
                if op == "||=" and truthy(cur):
                    return cur
                if op == "??=" and cur is not None and cur is not UNDEF:
                    return cur
                v = vf(env)
                setf(env, v)
                return v
            return assign_logical
        binop = self.BINOPS[op]
        if target[0] == "id":
            getf = self.getter(target[1], scope)
            setf = self.setter(target[1], scope)
            combine = self.e_bin(("bin", binop, ("__val", 0), ("__val", 1)), scope)

            def compound_id(env):
                v = combine((getf(env), vf(env)))
                setf(env, v)
                return v
            return compound_id
        objf, keyf = self.expr(target[1], scope), self.expr(target[2], scope)
        combine = self.e_bin(("bin", binop, ("__val", 0), ("__val", 1)), scope)

        def compound_member(env):
            o = objf(env)
            k = keyf(env)
            cur = get_prop(o, k)
            v = combine((cur, vf(env)))
            set_prop(o, k, v)
            return v
        return compound_member

    def e___val(self, e, scope):
        i = e[1]
        return lambda pair: pair[i]

    def e_update(self, e, scope):
        _, op, prefix, target = e
        delta = 1.0 if op == "++" else -1.0
        if target[0] == "id":
            getf = self.getter(target[1], scope)
            setf = self.setter(target[1], scope)
            hops = self.resolve(target[1], scope)
            name = target[1]
            if hops == 0 and not prefix:
                def post0(env):
                    v = env.v[name]
                    if type(v) is not float:
                        v = to_num(v)
                    env.v[name] = v + delta
                    return v
                return post0

            def upd_id(env):
                old = to_num(getf(env))
                setf(env, old + delta)
                return old + delta if prefix else old
            return upd_id
        objf, keyf = self.expr(target[1], scope), self.expr(target[2], scope)

        def upd_member(env):
            o = objf(env)
            k = keyf(env)
            old = to_num(get_prop(o, k))
            set_prop(o, k, old + delta)
            if prefix:
                return get_prop(o, k) if type(o) is JSTypedArray else old + delta
            return old
        return upd_member

    def e_member(self, e, scope):
        _, obj, prop, optional = e
        objf = self.expr(obj, scope)
        if prop[0] == "str":
            key = prop[1]
            if optional:
                def member_opt(env):
                    o = objf(env)
                    if o is None or o is UNDEF:
                        return UNDEF
                    return get_prop(o, key)
                return member_opt

            def member_named(env):
                o = objf(env)
                if type(o) is JSObject:
                    v = o.props.get(key, o)
                    if v is not o and type(v) is not Accessor:
                        return v
                return get_prop(o, key)
            return member_named
        keyf = self.expr(prop, scope)

        def member_computed(env):
            o = objf(env)
            k = keyf(env)
            if type(k) is float:
                t = type(o)
                if t is JSTypedArray:
                    i = int(k)
                    if i == k and 0 <= i < o.length:
                        return float(o.mv[i])
                    return UNDEF
                if t is JSArray:
                    i = int(k)
                    if i == k and 0 <= i < len(o):
                        return o[i]
                    return UNDEF
            if optional and (o is None or o is UNDEF):
                return UNDEF
            return get_prop(o, k)
        return member_computed

    def args(self, args, scope):
        if any(a[0] == "spread" for a in args):
            parts = [(a[0] == "spread", self.expr(a[1] if a[0] == "spread" else a, scope)) for a in args]

            def build(env):
                out = []
                for sp, f in parts:
                    if sp:
                        out.extend(iterate(f(env)))
                    else:
                        out.append(f(env))
                return out
            return build
        fs = [self.expr(a, scope) for a in args]
        n = len(fs)
        if n == 0:
            return lambda env: []
        if n == 1:
            f0 = fs[0]
            return lambda env: [f0(env)]
        if n == 2:
            f0, f1 = fs
            return lambda env: [f0(env), f1(env)]
        if n == 3:
            f0, f1, f2 = fs
            return lambda env: [f0(env), f1(env), f2(env)]
        return lambda env: [f(env) for f in fs]

    def e_call(self, e, scope):
        _, callee, args, optional, line = e
        argsf = self.args(args, scope)
        fname = os.path.basename(self.fname)
        if callee[0] == "member":
            objf = self.expr(callee[1], scope)
            prop = callee[2]
            member_optional = callee[3]
            key = prop[1] if prop[0] == "str" else None
            keyf = None if key is not None else self.expr(prop, scope)
            math_obj = self.globals.get("Math")

            def call_method(env):
                o = objf(env)
                k = key if key is not None else keyf(env)
                if (member_optional or optional) and (o is None or o is UNDEF):
                    return UNDEF
                t = type(o)
                if t is JSObject:
                    f = o.props.get(k, o) if type(k) is str else o
                    if f is o or type(f) is Accessor:
                        f = get_prop(o, k)
                elif t is JSArray:
                    m = ARRAY_METHODS.get(k)
                    if m is not None:
                        return _wrap(m(o, argsf(env)))
                    f = get_prop(o, k)
                elif t is JSTypedArray:
                    m = TYPED_METHODS.get(k)
                    if m is not None:
                        return _wrap(m(o, argsf(env)))
                    f = get_prop(o, k)
                elif t is JSSet:
                    m = SET_METHODS.get(k)
                    if m is not None:
                        return _wrap(m(o, argsf(env)))
                    f = UNDEF
                elif t is JSMap:
                    m = MAP_METHODS.get(k)
                    if m is not None:
                        return _wrap(m(o, argsf(env)))
                    f = UNDEF
                else:
                    f = get_prop(o, k, line)
                if optional and (f is None or f is UNDEF):
                    return UNDEF
                tf = type(f)
                if tf is HostFunction:
                    return _wrap(f.fn(o, argsf(env)))
                if tf is JSFunction and not f.is_class:
                    return f.call(o, argsf(env))
                throw_error("TypeError", f"{to_display(k)} is not a function ({fname}:{line})")
            return call_method
        if callee[0] == "super":
            raise NotImplementedError("super calls are not supported")
        ff = self.expr(callee, scope)

        def call_plain(env):
            f = ff(env)
            if optional and (f is None or f is UNDEF):
                return UNDEF
            tf = type(f)
            if tf is JSFunction and not f.is_class:
                return f.call(UNDEF, argsf(env))
            if tf is HostFunction:
                return _wrap(f.fn(UNDEF, argsf(env)))
            throw_error("TypeError", f"{to_display(callee[1] if callee[0] == 'id' else f)} is not a function ({fname}:{line})")
        return call_plain

    def e_new(self, e, scope):
        _, callee, args, line = e
        ff = self.expr(callee, scope)
        argsf = self.args(args, scope)
        return lambda env: construct(ff(env), argsf(env), line)

    def e_importmeta(self, e, scope):
        url = "file://" + self.fname
        return lambda env: JSObject(None, {"url": url})

    def e_spread(self, e, scope):
        raise NotImplementedError("spread outside of a call / array / object literal")


def _s32(n):
    n &= 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


# ---------------------------------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------------------------------


class Module:
    def __init__(self, path):
        self.path = path
        self.env = None
        self.exports = {}          # exported name → local name


class Interpreter:
    def __init__(self, root: str, host_modules=None, random_fn=None, import_map=None):
        """import_map: {(basename of the importing file, specifier): replacement path} — swaps a module's imports without
        touching its source (the drop-in test loads the reference worker over bindings/node/planet_worker_shim.mjs this way)"""
        self.root = root
        self.import_map = import_map or {}
        self.console = []
        self.globals = make_globals(self.console, random_fn)
        self.modules = {}
        self.host_modules = host_modules or {}

    def host_function(self, fn, name=""):
        return HostFunction(fn, name)

    def load(self, path: str, source: str | None = None) -> Module:
        """evaluates a module once; `source` supplies the text for a virtual path (relative imports resolve against that path)"""
        path = os.path.realpath(os.path.join(self.root, path)) if not os.path.isabs(path) else os.path.realpath(path)
        if path in self.modules:
            return self.modules[path]
        mod = self.modules[path] = Module(path)
        src = source if source is not None else open(path, encoding="utf-8").read()
        ast = parse(src, path)
        body = ast[1]
        scope = CScope(None, True)
        names = hoisted_vars(body, [])
        scope.names.update(names)
        scope.names.update(lexical_names(body))
        imports = [s for s in body if s[0] == "import"]
        for s in imports:
            for _, local in s[1]:
                scope.names.add(local)
        scope.names.add("*default*")
        env = mod.env = Env({n: UNDEF for n in names}, None)
        # imports: evaluate dependencies first, then copy the bindings (the reference has no `export let` that is reassigned)
        for s in imports:
            spec = s[2]
            spec = self.import_map.get((os.path.basename(path), spec), spec)
            if spec in self.host_modules:
                exports = self.host_modules[spec]
                getter = exports.get
            else:
                dep = self.load(os.path.join(os.path.dirname(path), spec))
                getter = lambda name, dep=dep: dep.env.v[dep.exports[name]] if name in dep.exports else _missing(name, spec)       # noqa: E731
                exports = dep
            for imported, local in s[1]:
                if imported == "*":
                    ns = JSObject()
                    src_names = exports.keys() if isinstance(exports, dict) else exports.exports.keys()
                    for n in src_names:
                        ns.props[n] = getter(n)
                    env.v[local] = ns
                else:
                    env.v[local] = getter(imported)
        for s in body:
            if s[0] != "export":
                continue
            if s[1] == "decl":
                d = s[2]
                if d[0] == "var":
                    for target, _ in d[2]:
                        for n in pattern_names(target, []):
                            mod.exports[n] = n
                else:
                    mod.exports[d[1]] = d[1]
            elif s[1] == "default":
                mod.exports["default"] = s[2][1]
            elif s[1] == "defaultexpr":
                mod.exports["default"] = "*default*"
            elif s[1] == "specs":
                if s[3] is not None:
                    raise NotImplementedError("export … from")
                for local, exported in s[2]:
                    mod.exports[exported] = local
        comp = Compiler(self, path)
        run = comp.statements(body, scope, function_level=True)
        run(env)
        return mod

    def get_export(self, path, name):
        mod = self.load(path)
        return mod.env.v[mod.exports[name]]

    def call(self, f, *args, this=UNDEF):
        return call_function(f, this, list(args))


def _missing(name, spec):
    raise KeyError(f"module {spec} has no export {name}")


# ---------------------------------------------------------------------------------------------------------------------
# host ↔ JS conversion helpers for the vector generator
# ---------------------------------------------------------------------------------------------------------------------
def to_python(v):
    """JS value → Python: typed arrays → numpy arrays, Sets → lists (insertion order), objects → dicts"""
    import numpy as np
    t = type(v)
    if t is JSTypedArray:
        dt = {"f": np.float32, "d": np.float64, "i": np.int32, "I": np.uint32, "h": np.int16, "H": np.uint16, "b": np.int8, "B": np.uint8}[v.code]
        return np.frombuffer(v.mv.tobytes(), dt).copy()
    if t is JSArray:
        return [to_python(x) for x in v]
    if t is JSSet:
        return [to_python(x) for x in v.iterate()]
    if t is JSMap:
        return {to_python(k): to_python(x) for k, x in v.iterate()}
    if t is JSObject:
        return {k: to_python(get_prop(v, k)) for k in own_keys(v) if type(v.props.get(k)) not in (JSFunction, HostFunction)}
    if v is UNDEF:
        return None
    return v


def from_python(v):
    import numpy as np
    if isinstance(v, np.ndarray):
        kind = {np.dtype(np.float32): "Float32Array", np.dtype(np.float64): "Float64Array", np.dtype(np.int32): "Int32Array",
                np.dtype(np.uint32): "Uint32Array", np.dtype(np.uint8): "Uint8Array", np.dtype(np.int8): "Int8Array",
                np.dtype(np.uint16): "Uint16Array", np.dtype(np.int16): "Int16Array"}[v.dtype]
        base = array.array(_TYPECODES[kind], np.ascontiguousarray(v).tobytes())
        return JSTypedArray(kind, base=base, mv=memoryview(base))
    if isinstance(v, bool):
        return v
    if isinstance(v, (int, float, np.integer, np.floating)):
        return float(v)
    if isinstance(v, str) or v is None:
        return v
    if isinstance(v, (list, tuple)):
        return JSArray(from_python(x) for x in v)
    if isinstance(v, set):
        s = JSSet()
        for x in sorted(v):
            s.add(from_python(x))
        return s
    if isinstance(v, dict):
        return JSObject(None, {prop_key(from_python(k)) if not isinstance(k, str) else k: from_python(x) for k, x in v.items()})
    return v
