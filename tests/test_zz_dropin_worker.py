"""The drop-in claim, executed: the UNMODIFIED reference worker (/root/reference/js/planet-worker.js: handleGenerate, handleReapply,
handleComputeClimate, handleEditRecompute, its private runPostProcessing / computeTriangleElevations / buildClimateFields) runs with
its import lines redirected to bindings/node/planet_worker_shim.mjs, which forwards every stage function to the Node-API addon
(bindings/node/planet_b200_addon.cc) and through it to the C ABI.  Nothing of the reference's compute modules is loaded; the
replies must equal the ones the all-JavaScript reference worker produced (tests/golden/reference_*.npz: generate, reapply,
computeClimate, editRecompute, importHeightmap, single- and dual-layer planets) bit for bit.

No Node and no JavaScript runtime exist in this image: the JavaScript (worker + shim) is evaluated by tests/golden/minijs.py, the
addon by the in-process Node-API runtime tests/napi_host/.  Needs the reference tree, so it only runs in the build container."""
import os

import numpy as np
import pytest

REFERENCE_JS = "/root/reference/js"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "bindings", "node", "planet_worker_shim.mjs")
SWAPPED = ["./rng.js", "./simplex-noise.js", "./sphere-mesh.js", "./coarse-plates.js", "./plates.js", "./elevation.js", "./super-plates.js",
           "./terrain-post.js", "./wind.js", "./ocean.js", "./precipitation.js", "./temperature.js", "./koppen.js"]

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE_JS), reason="the reference tree is only present in the build container")


def _native_module(addon):
    """`createRequire(import.meta.url)('./build/Release/planet_b200_addon.node')` → an object whose methods call the addon; typed
    arrays cross as views of the same memory, like in Node"""
    from tests.golden import minijs as js
    from tests.napi_host.host import JsError, Uint8Clamped

    def to_py(v):
        if type(v) is js.JSTypedArray:
            dt = {"f": np.float32, "d": np.float64, "i": np.int32, "B": np.uint8}[v.code]
            return np.frombuffer(v.mv, dt)                   # shares the evaluator's storage: in-place stages write through
        if type(v) is js.JSObject:
            return {k: to_py(x) for k, x in v.props.items()}
        if v is js.UNDEF:
            return None
        return v

    def to_js(v):
        if isinstance(v, np.ndarray):
            return js.from_python(np.asarray(v).view(np.ndarray) if isinstance(v, Uint8Clamped) else v)
        if isinstance(v, dict):
            return js.JSObject(None, {k: to_js(x) for k, x in v.items()})
        if v is None:
            return js.UNDEF
        return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v

    native = js.JSObject()
    for name in addon.exports:
        def method(this, args, name=name):
            try:
                return to_js(addon.call(name, *[to_py(a) for a in args]))
            except JsError as e:
                js.throw_error("TypeError" if type(e).__name__ == "JsTypeError" else "Error", str(e))
        native.props[name] = js.HostFunction(method, name)
    require = js.HostFunction(lambda this, args: native, "require")
    return {"createRequire": js.HostFunction(lambda this, args: require, "createRequire")}


@pytest.mark.parametrize("name", ["A_600", "D_import_600", "E_single_layer_400", "B_2500"])
def test_reference_worker_over_the_shim_reproduces_its_own_replies(name):
    from tests.emul.build_emul import build
    from tests.golden import minijs as js
    from tests.napi_host.host import NapiHost
    from tests.test_zz_reference_vectors import check_reply, command_for, load
    addon = NapiHost(build(), "emu")
    addon.setOption("mesh_order", "delaunator")            # the vectors are in the reference's neighbour order
    imports = {("planet-worker.js", spec): SHIM for spec in SWAPPED}
    cdn = "https://cdn.jsdelivr.net/npm/delaunator@5.0.1/+esm"
    it = js.Interpreter(REFERENCE_JS, host_modules={"node:module": _native_module(addon), cdn: {"default": js.UNDEF}}, import_map=imports)
    posted = []
    worker_self = js.JSObject()
    worker_self.props["postMessage"] = js.HostFunction(lambda this, args: (posted.append(js.to_python(args[0])), js.UNDEF)[1], "postMessage")
    it.globals["self"] = worker_self
    it.load("planet-worker.js")
    loaded = {os.path.basename(p) for p in it.modules}
    assert loaded == {"planet-worker.js", "planet_worker_shim.mjs"}, f"a compute module of the reference was loaded: {loaded}"

    commands, replies = load(name)
    stats = dict(float_elements=0, float_differing=0, int_elements=0, worst=0.0)
    for i, (cmd, (rmeta, arrays)) in enumerate(zip(commands, replies)):
        del posted[:]
        msg = command_for(cmd)
        if "plateDensity" in msg:
            msg["plateDensity"] = {str(k): v for k, v in msg["plateDensity"].items()}
        js.call_function(worker_self.props["onmessage"], worker_self, [js.JSObject(None, {"data": js.from_python(msg)})])
        out = [m for m in posted if m.get("type") != "progress"]
        assert len(out) == 1 and out[0]["type"] == rmeta["type"], out[0].get("message", out)
        reply = out[0]
        for key in ("plateDensity", "plateDensityLand", "plateDensityOcean", "plateVec"):
            if isinstance(reply.get(key), dict):
                reply[key] = {int(k): v for k, v in reply[key].items()}
        check_reply("dropin " + name, i, reply, rmeta, arrays, stats)
    assert stats["float_differing"] == 0 and stats["float_elements"] > 20000
