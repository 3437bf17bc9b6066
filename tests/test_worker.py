"""The worker command layer (js/planet-worker.js:136-677) over the engine: `generate`, `reapply`, `editRecompute`,
`computeClimate` with the reference's message fields and reply keys, against the same commands composed from the
oracle's stage functions — every array of every reply bit for bit."""
import numpy as np
import pytest

from planet_heightmap_generation_b200.sphere import park_miller
from planet_heightmap_generation_b200.worker import PlanetWorker

GEN = dict(cmd="generate", N=3000, P=12, jitter=0.75, nMag=0.4, numContinents=3, smoothing=0.1, glacialErosion=0.5, hydraulicErosion=0.5,
           thermalErosion=0.1, ridgeSharpening=0.5, terrainWarp=0.75, continentSizeVariety=0.3, temperatureOffset=0.0,
           precipitationOffset=0.0, landCoverage=0.3, seed=4242, toggledIndices=[1])
SLIDER_KEYS = ("smoothing", "glacialErosion", "hydraulicErosion", "thermalErosion", "ridgeSharpening", "terrainWarp")


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    if a.dtype == np.float32:
        return bool((a.view(np.uint32) == b.view(np.uint32)).all())
    return bool((a == b).all())


class OracleWorker:
    """The handlers of js/planet-worker.js restated over oracle/ (test infrastructure)."""

    def __init__(self, oracle, mesh_order="canonical"):
        self.o = oracle
        self.mesh_order = mesh_order          # "delaunator": the reference's own neighbour order (oracle/delaunator_ref.py)

    def climate(self, elev, t_off, p_off, cover, recompute_wind=True):
        c = self.clim
        if recompute_wind:
            c.wind(elev, self.pio, self.r_plate, self.seed)
            c.ocean(elev)
        c.precipitation(elev, p_off, cover)
        c.temperature(elev, t_off)
        return c.koppen(elev)

    def elevate(self, nMag):
        o = self.o
        table = {s: dict(isOcean=s in self.pio, pole=tuple(self.vec[s]["pole"]), omega=self.vec[s]["omega"], density=self.dens[s]) for s in self.seeds}
        r_super, sp = (o.build_super_plates(self.mesh, self.r_plate, table) if self.P >= 8 else (None, None))
        self.oe.assign(self.r_plate, table, self.seeds, self.seed, nMag, self.seed, 5, r_super, sp)
        return self.oe.get("r_elevation")

    def generate(self, d):
        o = self.o
        self.seed, self.P = d["seed"], d["P"]
        self.mesh, self.xyz = o.build_sphere(d["N"], d["jitter"], self.seed, self.mesh_order)
        self.nd = o.neighbor_dist(self.mesh, self.xyz)
        cp = o.generate_coarse_plates(self.seed, d["P"], d["numContinents"], d.get("continentSizeVariety", 0.0), d.get("landCoverage", 0.3),
                                      mesh_order=self.mesh_order)
        self.r_plate = o.project_coarse_plates(self.mesh, self.xyz, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], self.seed, d["P"])
        self.seeds, self.vec = cp["coarsePlateSeeds"], cp["coarsePlateVec"]
        o.smooth_and_reconnect_plates(self.mesh, self.r_plate, self.seeds, 3)
        self.pio = set(cp["coarsePlateIsOcean"])
        for i in d.get("toggledIndices") or []:
            self.pio.symmetric_difference_update({self.seeds[i]})
        self.dens = {}
        for s in self.seeds:
            r = park_miller(s + 777, 2)
            self.dens[s] = float(3.0 + r[0] * 0.5) if s in self.pio else float(2.4 + r[1] * 0.5)
        self.oe = o.Elevation(self.mesh, self.xyz)
        self.clim = o.Climate(self.mesh, self.xyz)
        elev = self.elevate(d["nMag"])
        self.pre = elev.copy()
        delta, _ = o.run_post_processing(self.mesh, self.xyz, elev, {k: d[k] for k in SLIDER_KEYS}, self.nd, self.seed, self.oe.get("hotspot"))
        self.final = elev.copy()
        koppen = self.climate(elev, d.get("temperatureOffset", 0.0), d.get("precipitationOffset", 0.0), d.get("landCoverage", 0.3))
        return elev, delta, koppen

    def reapply(self, d, t_off, p_off, cover):
        elev = self.pre.copy()
        delta, _ = self.o.run_post_processing(self.mesh, self.xyz, elev, {k: d[k] for k in SLIDER_KEYS}, self.nd, self.seed, None)
        self.final = elev.copy()
        koppen = None if d.get("skipClimate") else self.climate(elev, t_off, p_off, cover)
        return elev, delta, koppen

    def edit(self, d, t_off, p_off, cover):
        self.pio = set(d["plateIsOcean"])
        self.dens = dict(d["plateDensity"])
        elev = self.elevate(d["nMag"])
        self.pre = elev.copy()
        delta, _ = self.o.run_post_processing(self.mesh, self.xyz, elev, {k: d[k] for k in SLIDER_KEYS}, self.nd, self.seed, self.oe.get("hotspot"))
        self.final = elev.copy()
        koppen = self.climate(elev, t_off, p_off, cover)
        return elev, delta, koppen


def test_worker_commands_match_the_reference_handlers(backend, oracle):
    w = PlanetWorker(lib=backend)
    ow = OracleWorker(oracle)
    assert w.onmessage({"cmd": "reapply"})["type"] == "error"                    # no retained state yet (:342)
    assert w.onmessage({"cmd": "nonsense"}) == {"type": "error", "message": "Unknown command: nonsense"}

    # generate
    r = w.onmessage(dict(GEN))
    assert r["type"] == "done", r
    elev, delta, koppen = ow.generate(GEN)
    assert r["numRegions"] == ow.mesh.numRegions and same(r["r_xyz"], ow.xyz)
    assert same(r["triangles"], ow.mesh.triangles) and same(r["halfedges"], ow.mesh.halfedges)
    assert same(r["r_plate"], ow.r_plate) and r["plateSeeds"] == ow.seeds
    assert set(r["plateIsOcean"]) == ow.pio and r["plateDensity"] == ow.dens
    assert set(r["originalPlateIsOcean"]) ^ set(r["plateIsOcean"]) == {ow.seeds[1]}          # toggledIndices=[1]
    assert same(r["prePostElev"], ow.pre) and same(r["r_elevation"], elev) and same(r["debugLayers"]["erosionDelta"], delta)
    assert same(r["r_stress"], ow.oe.get("r_stress"))
    assert same(r["debugLayers"]["koppen"], koppen)
    for k in ("r_wind_east_summer", "r_ocean_warmth_winter", "r_precip_summer", "r_temperature_winter", "itczLatsSummer"):
        assert same(r[k], ow.clim.get(k)), k
    t = r["triangles"].reshape(-1, 3)
    e64 = elev.astype(np.float64)
    assert same(r["t_elevation"], (((e64[t[:, 0]] + e64[t[:, 1]]) + e64[t[:, 2]]) / 3).astype(np.float32))
    assert {"_pipelineTiming", "_postTiming", "_timing", "_params", "t_xyz", "mountain_r", "coastline_r", "ocean_r"} <= set(r)
    # the main thread's exports of that reply (js/planet-mesh.js:1752): Satellite and Köppen maps as PNG
    from planet_heightmap_generation_b200 import planet_mesh as pm
    for etype, fname in (("biome", "orogen-satellite-%s.png"), ("koppen", "orogen-climate-%s.png")):
        name, png = w.exportMap(r, etype, 256)
        assert name == fname % r["seed"]
        want_px, _ = oracle.export_map(ow.mesh, ow.xyz, etype, 256, elev, koppen)
        assert (pm.decode_png(png) == want_px).all()

    # reapply with other sliders, climate skipped → cached wind dropped (:386-389)
    msg = dict(cmd="reapply", smoothing=0.3, glacialErosion=0.2, hydraulicErosion=0.7, thermalErosion=0.0, ridgeSharpening=0.2,
               terrainWarp=0.0, skipClimate=True)
    r = w.onmessage(msg)
    elev, delta, _ = ow.reapply(msg, 0.0, 0.0, 0.3)
    assert r["type"] == "reapplyDone" and r["skipClimate"] and r["r_wind_east_summer"] is None and r["windDebugLayers"] is None
    assert same(r["r_elevation"], elev) and same(r["erosionDelta"], delta)

    # deferred climate with new offsets (:579-677): wind and ocean are recomputed because the cache was dropped
    r = w.onmessage(dict(cmd="computeClimate", temperatureOffset=2.5, precipitationOffset=-0.2))
    koppen = ow.climate(ow.final, 2.5, -0.2, 0.3)
    assert r["type"] == "climateDone" and same(r["climateDebugLayers"]["koppen"], koppen)
    assert same(r["r_temperature_summer"], ow.clim.get("r_temperature_summer")) and same(r["r_precip_winter"], ow.clim.get("r_precip_winter"))
    # again with other offsets: the cached wind / ocean are reused, the offsets stick (getClimateParams, :104-110)
    r = w.onmessage(dict(cmd="computeClimate", precipitationOffset=0.4))
    koppen = ow.climate(ow.final, 2.5, 0.4, 0.3, recompute_wind=False)
    assert same(r["climateDebugLayers"]["koppen"], koppen) and same(r["r_precip_summer"], ow.clim.get("r_precip_summer"))

    # editRecompute: flip one plate, change its density, new noise magnitude
    pio = set(ow.pio)
    pio.symmetric_difference_update({ow.seeds[3]})
    dens = dict(ow.dens)
    dens[ow.seeds[3]] = 2.95
    msg = dict(cmd="editRecompute", plateIsOcean=sorted(pio), plateDensity=dens, nMag=0.25, smoothing=0.1, glacialErosion=0.5,
               hydraulicErosion=0.5, thermalErosion=0.1, ridgeSharpening=0.5, terrainWarp=0.75)
    r = w.onmessage(msg)
    elev, delta, koppen = ow.edit(msg, 2.5, 0.4, 0.3)
    assert r["type"] == "editDone", r
    assert same(r["prePostElev"], ow.pre) and same(r["r_elevation"], elev) and same(r["debugLayers"]["erosionDelta"], delta)
    assert same(r["debugLayers"]["koppen"], koppen) and same(r["r_stress"], ow.oe.get("r_stress"))
    w.close()


def test_import_heightmap_command(backend, oracle):
    """importHeightmap (:771-942): bilinear sampling of an equirectangular grayscale image, post-processing, synthetic plates
    from the land / ocean components, region classes, climate — against the oracle's restatement of the same handler."""
    from oracle.mesh_hull import build_sphere_from_points
    rng = np.random.default_rng(3)
    W_, H_ = 96, 48
    yy, xx = np.mgrid[0:H_, 0:W_]
    img = 110 + 90 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 40 * rng.random((H_, W_))
    img[(np.sin(xx / 5.0 + yy / 11.0) > 0.2)] = 0            # oceans: black
    gray = np.clip(np.round(img), 0, 255).astype(np.uint8)
    msg = dict(cmd="importHeightmap", N=3000, jitter=0.75, grayscale=gray.reshape(-1), imageWidth=W_, imageHeight=H_, smoothing=0.1,
               glacialErosion=0.5, hydraulicErosion=0.5, thermalErosion=0.1, ridgeSharpening=0.5, terrainWarp=0.75, seed=77)
    w = PlanetWorker(lib=backend)
    r = w.onmessage(msg)
    assert r["type"] == "done", r
    mesh, xyz = build_sphere_from_points(oracle.fibonacci_sphere(3000, 0.75, 77))
    nd = oracle.neighbor_dist(mesh, xyz)
    elev = oracle.sample_heightmap(mesh, xyz, gray.reshape(-1), W_, H_)
    assert same(r["prePostElev"], elev) and (elev == np.float32(-0.5)).any() and (elev > 0).any()
    delta, _ = oracle.run_post_processing(mesh, xyz, elev, {k: msg[k] for k in SLIDER_KEYS}, nd, 77, None)
    assert same(r["r_elevation"], elev) and same(r["debugLayers"]["erosionDelta"], delta)
    r_plate, seeds, pio = oracle.derive_synthetic_plates(mesh, elev)
    assert same(r["r_plate"], r_plate) and r["plateSeeds"] == seeds and set(r["plateIsOcean"]) == pio and len(seeds) >= 2
    m, c, o = oracle.classify_imported(mesh, elev)
    assert r["mountain_r"] == [int(i) for i in np.nonzero(m)[0]] and r["coastline_r"] == [int(i) for i in np.nonzero(c)[0]]
    assert r["ocean_r"] == [int(i) for i in np.nonzero(o)[0]] and not r["r_stress"].any()
    clim = oracle.Climate(mesh, xyz)
    koppen = clim.run_all(elev, pio, r_plate, 77)
    assert same(r["debugLayers"]["koppen"], koppen) and same(r["r_precip_summer"], clim.get("r_precip_summer"))
    # the imported planet is retained: reapply works on it (:341)
    r2 = w.onmessage(dict(cmd="reapply", smoothing=0.0, glacialErosion=0.0, hydraulicErosion=0.3, thermalErosion=0.0, ridgeSharpening=0.0,
                          terrainWarp=0.0, skipClimate=True))
    elev2 = oracle.sample_heightmap(mesh, xyz, gray.reshape(-1), W_, H_)
    oracle.run_post_processing(mesh, xyz, elev2, dict(smoothing=0.0, glacialErosion=0.0, hydraulicErosion=0.3, thermalErosion=0.0,
                                                      ridgeSharpening=0.0, terrainWarp=0.0), nd, 77, None)
    assert r2["type"] == "reapplyDone" and same(r2["r_elevation"], elev2)
    w.close()


def test_failed_generate_keeps_previous_planet(emu_lib):
    """js/planet-worker.js:277 assigns W only after a successful run: a failing generate must leave the retained planet usable."""
    from planet_heightmap_generation_b200.worker import PlanetWorker
    w = PlanetWorker(lib=emu_lib)
    ok = w.onmessage(dict(cmd="generate", N=2000, P=8, jitter=0.75, nMag=0.4, numContinents=3, seed=7, skipClimate=True, **{k: 0.3 for k in SLIDER_KEYS}))
    assert ok["type"] == "done"
    bad = w.onmessage(dict(cmd="generate", N=2000, P=8, jitter=0.75, nMag=0.4, numContinents=3, seed=None, **{k: 0.3 for k in SLIDER_KEYS}))
    assert bad["type"] == "error"
    bad = w.onmessage(dict(cmd="generate", N=-5, P=8, jitter=0.75, nMag=0.4, numContinents=3, seed=3, **{k: 0.3 for k in SLIDER_KEYS}))
    assert bad["type"] == "error"
    again = w.onmessage(dict(cmd="reapply", skipClimate=True, **{k: 0.5 for k in SLIDER_KEYS}))
    assert again["type"] == "reapplyDone", again
    w.close()
