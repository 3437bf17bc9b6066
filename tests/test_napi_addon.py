"""The Node-API addon (bindings/node/planet_b200_addon.cc) EXECUTED — registration, argument unpacking, typed-array checks,
result objects, exceptions — by a minimal in-process Node-API runtime (tests/napi_host/, no Node in this image), linked
against the host emulation of the C ABI (CPU suite) or libplanet_b200.so (-m gpu).  The calls are the ones
bindings/node/planet_worker_shim.mjs makes for js/planet-worker.js's handleGenerate (:136-339); every array is compared bit
for bit with the oracle."""
import numpy as np
import pytest

from tests.conftest import _has_cuda, assert_bit_equal
from tests.napi_host.host import JsError, JsTypeError, NapiHost, Uint8Clamped

SLIDERS = dict(smoothing=0.10, glacialErosion=0.50, hydraulicErosion=0.50, thermalErosion=0.10, ridgeSharpening=0.50, terrainWarp=0.75)
N, P, CONT, SEED = 3000, 24, 3, 42.0


@pytest.fixture(scope="module", params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def addon(request):
    if request.param == "emu":
        from tests.emul.build_emul import build
        return NapiHost(build(), "emu")
    if not _has_cuda():
        pytest.fail("gpu test selected but no CUDA device is visible")
    import os
    from planet_heightmap_generation_b200 import build as b
    return NapiHost(b.SO if os.path.exists(b.SO) else b.build(), "cuda")


def test_addon_exports(addon):
    want = {"setMesh", "setOption", "computeNeighborDist", "warpTerrain", "smoothElevation", "erodeComposite", "sharpenRidges",
            "applySoilCreep", "runPostProcessing", "assignElevationFlat", "computeWindFlat", "computeOceanCurrentsFlat",
            "computePrecipitationFlat", "computeTemperatureFlat", "classifyKoppenFlat", "getClimateField", "smoothField",
            "exportMapPixels", "getMeshTriangles", "generateTriangleCenters", "buildSphereFlat", "generateCoarsePlatesFlat", "projectCoarsePlatesFlat",
            "smoothAndReconnectPlatesFlat", "buildSuperPlatesFlat"}
    assert set(addon.exports) == want


def test_addon_generate_chain_matches_oracle(addon, oracle):
    from oracle.mesh_hull import build_sphere_from_points
    from planet_heightmap_generation_b200.sphere import park_miller

    # buildSphere (:149)
    s = addon.buildSphereFlat(N, 0.75, SEED)
    mesh, xyz = build_sphere_from_points(oracle.fibonacci_sphere(N, 0.75, SEED))
    assert int(s["numRegions"]) == mesh.numRegions
    assert_bit_equal(s["r_xyz"], xyz, "r_xyz")
    assert_bit_equal(s["adjOffset"], mesh.adjOffset, "adjOffset")
    assert_bit_equal(s["adjList"], mesh.adjList, "adjList")
    nd = addon.computeNeighborDist()
    assert_bit_equal(nd, oracle.neighbor_dist(mesh, xyz), "neighborDist")
    tri = addon.getMeshTriangles()
    assert_bit_equal(tri["triangles"], mesh.triangles, "triangles")
    assert_bit_equal(tri["halfedges"], mesh.halfedges, "halfedges")
    assert_bit_equal(addon.generateTriangleCenters(), oracle.triangle_centers(mesh, xyz), "t_xyz")

    # generateCoarsePlates → projectCoarsePlates → smoothAndReconnectPlates (:160-173)
    cp = addon.generateCoarsePlatesFlat(SEED, P, CONT, 0.0, 0.3)
    ocp = oracle.generate_coarse_plates(SEED, P, CONT, 0.0, 0.3)
    k = int(cp["numPlates"])
    seeds = [int(v) for v in cp["seeds"][:k]]
    assert seeds == ocp["coarsePlateSeeds"]
    assert {sd for i, sd in enumerate(seeds) if cp["isOcean"][i]} == ocp["coarsePlateIsOcean"]
    assert_bit_equal(cp["coarse_r_plate"], ocp["coarse_r_plate"], "coarse_r_plate")
    r_plate = addon.projectCoarsePlatesFlat(int(cp["numRegions"]), cp["adjOffset"], cp["adjList"], cp["coarse_xyz"], cp["coarse_r_plate"], SEED, P)
    seeds_a = np.asarray(seeds, np.int32)
    addon.smoothAndReconnectPlatesFlat(r_plate, seeds_a, 3)                 # in place, like the reference
    o_plate = oracle.project_coarse_plates(mesh, xyz, ocp["coarseMesh"], ocp["coarse_xyz"], ocp["coarse_r_plate"], SEED, P)
    oracle.smooth_and_reconnect_plates(mesh, o_plate, seeds, 3)
    assert_bit_equal(r_plate, o_plate, "r_plate")

    # plate tables with the worker's densities (:193-201), buildSuperPlates (:209)
    pio = ocp["coarsePlateIsOcean"]
    dens = {sd: float(3.0 + park_miller(sd + 777, 2)[0] * 0.5) if sd in pio else float(2.4 + park_miller(sd + 777, 2)[1] * 0.5) for sd in seeds}
    ids, oc = seeds_a, np.ascontiguousarray(cp["isOcean"][:k])
    pole, omega = np.ascontiguousarray(cp["pole"][:3 * k]), np.ascontiguousarray(cp["omega"][:k])
    density = np.asarray([dens[sd] for sd in seeds])
    assert_bit_equal(np.ascontiguousarray(cp["density"][:k]), density, "plateDensity")
    sp = addon.buildSuperPlatesFlat(r_plate, ids, oc, pole, omega, density)
    plates = {sd: dict(isOcean=sd in pio, pole=tuple(ocp["coarsePlateVec"][sd]["pole"]), omega=ocp["coarsePlateVec"][sd]["omega"], density=dens[sd]) for sd in seeds}
    o_super, o_sp = oracle.build_super_plates(mesh, r_plate, plates)
    assert_bit_equal(sp["r_superPlate"], o_super, "r_superPlate")
    ns = int(sp["numSuperPlates"])
    assert ns == len(o_sp)

    # assignElevation (:215), dual layer
    s_ids = np.arange(ns, dtype=np.int32)
    res = addon.assignElevationFlat(r_plate, ids, oc, pole, omega, density, seeds_a, SEED, 0.4, SEED, 5,
                                    sp["r_superPlate"], s_ids, np.ascontiguousarray(sp["isOcean"][:ns]), np.ascontiguousarray(sp["pole"][:3 * ns]),
                                    np.ascontiguousarray(sp["omega"][:ns]), np.ascontiguousarray(sp["density"][:ns]))
    oe = oracle.Elevation(mesh, xyz)
    oe.assign(o_plate, plates, seeds, SEED, 0.4, SEED, 5, o_super, o_sp)
    assert_bit_equal(res["r_elevation"], oe.get("r_elevation"), "r_elevation")
    assert_bit_equal(res["r_stress"], oe.get("r_stress"), "r_stress")
    for key in ("mountain_r", "coastline_r", "ocean_r"):
        assert_bit_equal(res[key], oe.get(key, np.uint8), key)
    assert_bit_equal(res["debugLayers"]["hotspot"], oe.get("hotspot"), "debugLayers.hotspot")

    # runPostProcessing (:223): r_elevation mutated in place, {dl_erosionDelta} returned
    elev = res["r_elevation"].copy()
    want = oe.get("r_elevation")
    o_delta, o_ocean = oracle.run_post_processing(mesh, xyz, want, SLIDERS, nd, SEED, oe.get("hotspot"))
    post = addon.runPostProcessing(None, s["r_xyz"], elev, SLIDERS, nd, SEED, res["debugLayers"]["hotspot"])
    assert_bit_equal(elev, want, "runPostProcessing elevation")
    assert_bit_equal(post["dl_erosionDelta"], o_delta, "dl_erosionDelta")
    assert set(post["postTimingMs"]) == {"Terrain warp", "Smoothing", "Erosion composite", "Ridge sharpening", "Soil creep"}

    # climate (:232-266) and result-object field reads
    ocn = oracle.Climate(mesh, xyz)
    o_koppen = ocn.run_all(want, pio, o_plate, SEED)
    addon.computeWindFlat(elev, np.asarray(sorted(pio), np.int32), r_plate, SEED, 23.5)
    addon.computeOceanCurrentsFlat(elev)
    addon.computePrecipitationFlat(elev, 0.0, 0.3)
    addon.computeTemperatureFlat(elev, 0.0)
    koppen = addon.classifyKoppenFlat(elev)
    assert_bit_equal(koppen, o_koppen, "r_koppen")
    for key in ("r_wind_east_summer", "r_ocean_warmth_winter", "r_precip_summer", "r_temperature_winter"):
        assert_bit_equal(addon.getClimateField(key), ocn.get(key), key)
    assert_bit_equal(addon.getClimateField("r_coastDistLand"), ocn.get("r_coastDistLand", np.int32), "r_coastDistLand")

    # main-thread export (js/planet-mesh.js:1752): Uint8ClampedArray for `new ImageData(px, width)`
    px = addon.exportMapPixels("biome", 256, elev, koppen)
    assert isinstance(px, Uint8Clamped)
    o_px, _ = oracle.export_map(mesh, xyz, "biome", 256, want, o_koppen)
    assert (px.reshape(128, 256, 4) == o_px).all()
    assert (addon.exportMapPixels("koppen", 128, elev, None) == addon.exportMapPixels("colormap", 128, elev, None)).all()


def test_addon_stage_functions_in_place(addon, oracle, planet_small):
    """setMesh with a caller-built mesh (the path that keeps the reference's own neighbour order), then the five terrain-post
    stage functions one by one on the caller's Float32Array."""
    mesh, xyz, nd, elev = planet_small()
    addon.setMesh(mesh.numRegions, mesh.adjOffset, mesh.adjList, xyz)
    want = elev.copy()
    hot = np.zeros_like(elev)
    oracle.warp_terrain(mesh, want, xyz, 42, 0.5, hot)
    addon.warpTerrain(None, elev, xyz, 42, 0.5, hot)
    assert_bit_equal(elev, want, "warpTerrain")
    is_ocean = (elev <= 0).astype(np.uint8)
    oracle.smooth_elevation(mesh, want, is_ocean, 2, 0.4)
    addon.smoothElevation(None, elev, is_ocean, 2, 0.4)
    assert_bit_equal(elev, want, "smoothElevation")
    oracle.erode_composite(mesh, want, xyz, is_ocean, 4, 0.0003, 0.5, 1, 1, 1.16, 0.015, 2, 0.5, nd)
    addon.erodeComposite(None, elev, xyz, is_ocean, 4, 0.0003, 0.5, 1, 1, 1.16, 0.015, 2, 0.5, nd)
    assert_bit_equal(elev, want, "erodeComposite")
    oracle.sharpen_ridges(mesh, want, is_ocean, 3, 0.04)
    addon.sharpenRidges(None, elev, is_ocean, 3, 0.04)
    oracle.apply_soil_creep(mesh, want, is_ocean, 3, 0.1125)
    addon.applySoilCreep(None, elev, is_ocean, 3, 0.1125)
    assert_bit_equal(elev, want, "sharpenRidges + applySoilCreep")
    f = elev.copy()
    oracle.smooth_field(mesh, want, 3)
    addon.smoothField(None, f, 3)
    assert_bit_equal(f, want, "smoothField")


def test_addon_rejects_bad_arguments(addon, oracle, planet_small):
    mesh, xyz, nd, elev = planet_small()
    addon.setMesh(mesh.numRegions, mesh.adjOffset, mesh.adjList, xyz)
    is_ocean = (elev <= 0).astype(np.uint8)
    with pytest.raises(JsTypeError, match="wrong element type"):
        addon.smoothElevation(None, elev.astype(np.float64), is_ocean, 1, 0.4)          # Float64Array where Float32Array is read
    with pytest.raises(JsTypeError, match="wrong length"):
        addon.smoothElevation(None, elev[:-1].copy(), is_ocean, 1, 0.4)                  # stale array from another mesh
    with pytest.raises(JsTypeError, match="not a typed array"):
        addon.smoothElevation(None, {"length": 3}, is_ocean, 1, 0.4)
    with pytest.raises(JsTypeError, match="adjList length"):
        addon.setMesh(mesh.numRegions, mesh.adjOffset, mesh.adjList[:-1].copy(), xyz)
    with pytest.raises(JsError, match="setMesh"):                                        # the failed setMesh dropped the mesh
        addon.smoothField(None, elev, 1)
    addon.setMesh(mesh.numRegions, mesh.adjOffset, mesh.adjList, xyz)
    with pytest.raises(JsTypeError, match="width"):
        addon.exportMapPixels("colormap", 333, elev, None)
    with pytest.raises(JsError, match="unknown option"):                                 # pb_last_error() surfaces as Error(message)
        addon.setOption("no_such_option", "x")
    before = elev.copy()
    with pytest.raises(JsTypeError):
        addon.warpTerrain(None, elev, xyz, 42, 0.5, elev.astype(np.float64))
    assert_bit_equal(elev, before, "a rejected call must not touch its arguments")
