"""Host mirror of the reference's Web Worker command layer (js/planet-worker.js:136-677, 944-954).

`PlanetWorker.onmessage(data)` takes the same message objects the main thread posts (`cmd`: generate | reapply |
editRecompute | computeClimate | importHeightmap, same field names) and returns the reply the worker would post (`type`: done |
reapplyDone | editDone | climateDone | error, same keys), with typed arrays as numpy arrays and JS Sets posted as lists.
`importHeightmap` (:771-942) is handled too.  Every stage runs through the C ABI on the GPU; the retained state `W`
(:277-292) is one DeviceMesh (mesh, r_xyz, neighborDist and the climate fields stay in HBM) plus the small host-side
plate tables.

Differences from the reference, by construction:
  * `seed` must be given (the reference draws `Math.random()` when it is missing, :145);
  * `_timing` / `_pipelineTiming` / `_postTiming` carry wall-clock ms of these stages, not of the JS ones;
  * progress messages are handed to an optional callback instead of being posted;
  * `mountain_r` / `coastline_r` / `ocean_r` are posted in ascending id (the C ABI returns membership masks), not in the
    Sets' insertion order — the main thread only tests membership and draws the regions.
"""
from __future__ import annotations

import time

import numpy as np

from . import climate as cl
from . import plates as pl
from .elevation import assignElevation
from .engine import DeviceMesh
from .sphere import park_miller
from .terrain_post import runPostProcessing

_SLIDERS = ("smoothing", "glacialErosion", "hydraulicErosion", "thermalErosion", "ridgeSharpening", "terrainWarp")
_WIND = ("r_wind_east_summer", "r_wind_north_summer", "r_wind_east_winter", "r_wind_north_winter", "itczLons", "itczLatsSummer", "itczLatsWinter")
_OCEAN = tuple(f"r_ocean_{k}_{s}" for k in ("current_east", "current_north") for s in ("summer", "winter")) + \
    ("r_ocean_speed_summer", "r_ocean_speed_winter", "r_ocean_warmth_summer", "r_ocean_warmth_winter")
_PRECIP = ("r_precip_summer", "r_precip_winter")
_TEMP = ("r_temperature_summer", "r_temperature_winter")


def _mask_to_list(mask) -> list:
    return [int(i) for i in np.nonzero(np.asarray(mask))[0]]


class PlanetWorker:
    def __init__(self, device: int = 0, lib=None, progress=None, mesh_order: str = "canonical"):
        """mesh_order: "canonical" (device mesh builder) or "delaunator" (the reference's own neighbour order — same seed,
        same planet as the web app; host algorithm, see DeviceMesh.build_sphere)."""
        self.device, self.lib, self.progress = device, lib, progress or (lambda pct, label: None)
        self.mesh_order = mesh_order
        self.W = None
        self._building = None

    # ---- dispatch (js/planet-worker.js:944-954) ----------------------------------------------------------------------
    def onmessage(self, data: dict) -> dict:
        handlers = {"generate": self._generate, "reapply": self._reapply, "editRecompute": self._edit_recompute,
                    "computeClimate": self._compute_climate, "importHeightmap": self._import_heightmap}
        cmd = data.get("cmd")
        if cmd not in handlers:
            return {"type": "error", "message": f"Unknown command: {cmd}"}
        if cmd not in ("generate", "importHeightmap") and self.W is None:
            return {"type": "error", "message": f"No retained state for {cmd}"}
        self._building = None
        try:
            return handlers[cmd](data)
        except Exception as e:      # the handlers' try/catch (:336-338)
            # the reference assigns W only after a successful run (:277): a failed generate / import leaves the previous planet
            # usable for reapply / editRecompute / computeClimate; the half-built mesh is released
            if self._building is not None:
                try:
                    self._building.close()
                except Exception:
                    pass
            return {"type": "error", "message": str(e)}
        finally:
            self._building = None

    def close(self):
        if self.W is not None:
            self.W["mesh"].close()
            self.W = None

    def exportMap(self, curData: dict, type: str, width: int):
        """exportMap(type, width) (js/planet-mesh.js:1752-1961).  In the reference this runs on the main thread over
        `state.curData` — the last done / reapplyDone / editDone reply (its `r_elevation`, `debugLayers.koppen`, `seed`) — and
        the mesh; here the mesh is the retained one in HBM.  → (download filename, PNG bytes)"""
        from . import planet_mesh as pm
        if self.W is None:
            raise RuntimeError("No retained state for exportMap")
        layers = curData.get("debugLayers") or {}
        return pm.exportMap(self.W["mesh"], type, width, curData["r_elevation"], layers.get("koppen"), seed=curData.get("seed", ""))

    def _retain(self, new_state):
        """W = new_state once the run has succeeded (:277-292); the planet held before is released"""
        old, self.W = self.W, new_state
        self._building = None
        if old is not None and old["mesh"] is not new_state["mesh"]:
            old["mesh"].close()

    # ---- shared pieces ---------------------------------------------------------------------------------------------------
    def _climate_params(self, data):          # getClimateParams (:104-110)
        W = self.W or {}
        out = {k: data.get(k) if data.get(k) is not None else W.get(k, d)
               for k, d in (("temperatureOffset", 0), ("precipitationOffset", 0), ("landCoverage", 0.3))}
        if self.W is not None:
            self.W.update(out)
        return out["temperatureOffset"], out["precipitationOffset"], out["landCoverage"]

    @staticmethod
    def _climate_fields(wind, ocean, precip, temp):       # buildClimateFields (:112-134)
        out = {}
        for res, keys in ((wind, _WIND), (ocean, _OCEAN), (precip, _PRECIP), (temp, _TEMP)):
            for k in keys:
                out[k] = None if res is None else res[k]
        return out

    def _climate(self, mesh, elev, pio, r_plate, seed, temperatureOffset, precipitationOffset, landCoverage, layers=None):
        """wind → ocean → precipitation → temperature → Köppen (:229-266); fills the debug layers the handler names."""
        t = {}
        t0 = time.perf_counter()
        wind = cl.computeWind(mesh, mesh.r_xyz, elev, pio, r_plate, seed)
        t["wind"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        ocean = cl.computeOceanCurrents(mesh, mesh.r_xyz, elev, wind)
        t["ocean"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        precip = cl.computePrecipitation(mesh, mesh.r_xyz, elev, wind, ocean, precipitationOffset, landCoverage)
        t["precipitation"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        temp = cl.computeTemperature(mesh, mesh.r_xyz, elev, wind, ocean, precip, temperatureOffset)
        t["temperature"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        koppen = cl.classifyKoppen(mesh, elev, temp, precip)
        t["koppen"] = 1e3 * (time.perf_counter() - t0)
        if layers is not None:
            layers.update(pressureSummer=wind["r_pressure_summer"], pressureWinter=wind["r_pressure_winter"],
                          windSpeedSummer=wind["r_wind_speed_summer"], windSpeedWinter=wind["r_wind_speed_winter"],
                          continentality=wind["r_continentality"],
                          precipSummer=precip["r_precip_summer"], precipWinter=precip["r_precip_winter"],
                          rainShadowSummer=precip["r_rainshadow_summer"], rainShadowWinter=precip["r_rainshadow_winter"],
                          tempSummer=temp["r_temperature_summer"], tempWinter=temp["r_temperature_winter"], koppen=koppen)
        return wind, ocean, precip, temp, koppen, t

    @staticmethod
    def _densities(seeds, pio):                # js/planet-worker.js:193-201
        land, ocean, dens = {}, {}, {}
        for r in seeds:
            d = park_miller(r + 777, 2)
            ocean[r] = float(3.0 + d[0] * 0.5)
            land[r] = float(2.4 + d[1] * 0.5)
            dens[r] = ocean[r] if r in pio else land[r]
        return dens, land, ocean

    # ---- generate (:136-339) ---------------------------------------------------------------------------------------------
    def _generate(self, data):
        N, P, jitter, nMag = int(data["N"]), int(data["P"]), float(data["jitter"]), float(data["nMag"])
        numContinents = int(data["numContinents"])
        variety = float(data.get("continentSizeVariety", 0) or 0)
        skip = bool(data.get("skipClimate"))
        if data.get("seed") is None:
            raise ValueError("generate needs a seed (the reference would draw Math.random())")
        seed = data["seed"]
        temperatureOffset, precipitationOffset, landCoverage = self._climate_params(data)
        timing = []

        def stage(name, t0):
            timing.append({"stage": name, "ms": 1e3 * (time.perf_counter() - t0)})

        self.progress(0, "Shaping the world…")
        t0 = time.perf_counter()
        mesh = self._building = DeviceMesh.build_sphere(N, jitter, seed, device=self.device, lib=self.lib, order=self.mesh_order)
        stage("Sphere mesh (Fibonacci + Delaunay + pole)", t0); t0 = time.perf_counter()
        neighborDist = mesh.computeNeighborDist()
        stage("Neighbor distances", t0); t0 = time.perf_counter()
        t_xyz = mesh.generateTriangleCenters()
        triangles, halfedges = mesh.trianglesAndHalfedges()
        stage("Triangle centers", t0); t0 = time.perf_counter()
        self.progress(10, "Generating coarse plates…")
        cp = pl.generateCoarsePlates(mesh, seed, P, numContinents, variety, landCoverage)
        stage(f"Coarse plates ({P} plates, {numContinents} continents)", t0); t0 = time.perf_counter()
        self.progress(20, "Projecting plates…")
        r_plate = pl.projectCoarsePlates(mesh, mesh.r_xyz, cp["coarseMesh"], cp["coarse_xyz"], cp["coarse_r_plate"], seed, P)
        cp["coarseMesh"].close()
        stage("Project coarse → hi-res", t0); t0 = time.perf_counter()
        self.progress(25, "Smoothing boundaries…")
        plateSeeds, plateVec = cp["coarsePlateSeeds"], cp["coarsePlateVec"]
        pl.smoothAndReconnectPlates(mesh, r_plate, plateSeeds, 3)
        stage("Smooth projected plates", t0)
        plateIsOcean = set(cp["coarsePlateIsOcean"])
        original = set(plateIsOcean)
        for i in data.get("toggledIndices") or []:            # :181-190
            if i < len(plateSeeds):
                r = plateSeeds[i]
                plateIsOcean.symmetric_difference_update({r})
        plateDensity, plateDensityLand, plateDensityOcean = self._densities(plateSeeds, plateIsOcean)
        superPlateData = None
        if P >= 8:
            t0 = time.perf_counter()
            superPlateData = pl.buildSuperPlates(mesh, r_plate, plateSeeds, plateVec, plateIsOcean, plateDensity)
            stage(f"Super plates ({superPlateData['numSuperPlates']} groups from {P} plates)", t0)
        self.progress(35, "Raising mountains…")
        t0 = time.perf_counter()
        res = assignElevation(mesh, mesh.r_xyz, plateIsOcean, r_plate, plateVec, plateSeeds, seed, nMag, seed, 5, plateDensity, superPlateData)
        stage("Elevation (collisions + stress + distance fields + assignment)", t0)
        r_elevation, debugLayers = res["r_elevation"], res["debugLayers"]
        if superPlateData is not None:
            debugLayers["superPlates"] = np.asarray(superPlateData["r_superPlate"], np.float32)      # js/elevation.js:1384
        prePostElev = r_elevation.copy()
        self.progress(60, "Eroding terrain…")
        t0 = time.perf_counter()
        post = runPostProcessing(mesh, mesh.r_xyz, r_elevation, {k: data.get(k, 0) or 0 for k in _SLIDERS}, neighborDist, seed,
                                 debugLayers["hotspot"])
        stage("Terrain post-processing (total)", t0)
        debugLayers["erosionDelta"] = post["dl_erosionDelta"]
        wind = ocean = precip = temp = None
        if not skip:
            self.progress(70, "Simulating wind patterns…")
            t0 = time.perf_counter()
            wind, ocean, precip, temp, _, _ = self._climate(mesh, r_elevation, plateIsOcean, r_plate, seed, temperatureOffset,
                                                            precipitationOffset, landCoverage, debugLayers)
            stage("Climate (wind, ocean currents, precipitation, temperature, Köppen)", t0)
        self.progress(75 if skip else 90, "Computing triangle elevations…")
        t0 = time.perf_counter()
        t_elevation = mesh.computeTriangleElevations(r_elevation)
        stage("Triangle elevations", t0)
        self._retain(dict(mesh=mesh, neighborDist=neighborDist, r_plate=r_plate.copy(), plateSeeds=list(plateSeeds), plateVec=plateVec,
                      plateIsOcean=set(plateIsOcean), originalPlateIsOcean=set(original), plateDensity=dict(plateDensity),
                      plateDensityLand=plateDensityLand, plateDensityOcean=plateDensityOcean, prePostElev=prePostElev.copy(),
                      r_elevation_final=r_elevation.copy(), seed=seed, nMag=nMag, P=P,
                      temperatureOffset=temperatureOffset, precipitationOffset=precipitationOffset, landCoverage=landCoverage,
                      cachedWind=wind, cachedOcean=ocean))
        reply = {"type": "done", "triangles": triangles, "halfedges": halfedges, "numRegions": mesh.numRegions,
                 "r_xyz": mesh.r_xyz, "t_xyz": t_xyz, "r_plate": r_plate, "plateSeeds": list(plateSeeds), "plateVec": plateVec,
                 "plateIsOcean": [s for s in plateSeeds if s in plateIsOcean] + [s for s in plateIsOcean if s not in plateSeeds],
                 "originalPlateIsOcean": [s for s in plateSeeds if s in original],
                 "plateDensity": plateDensity, "plateDensityLand": plateDensityLand, "plateDensityOcean": plateDensityOcean,
                 "prePostElev": prePostElev, "r_elevation": r_elevation, "t_elevation": t_elevation,
                 "mountain_r": _mask_to_list(res["mountain_r"]), "coastline_r": _mask_to_list(res["coastline_r"]),
                 "ocean_r": _mask_to_list(res["ocean_r"]), "r_stress": res["r_stress"]}
        reply.update(self._climate_fields(wind, ocean, precip, temp))
        reply.update(skipClimate=skip, seed=seed, nMag=nMag, debugLayers=debugLayers, _timing=res.get("_timing", []),
                     _pipelineTiming=timing, _postTiming=post.get("postTiming", []), _workerTotal=sum(s["ms"] for s in timing),
                     _params={k: data.get(k) for k in ("N", "P", "jitter", "nMag", "numContinents", "smoothing", "terrainWarp",
                                                       "hydraulicErosion", "thermalErosion", "ridgeSharpening", "glacialErosion",
                                                       "continentSizeVariety", "temperatureOffset", "precipitationOffset",
                                                       "landCoverage")} | {"seed": seed})
        return reply

    # ---- importHeightmap (:771-942) ------------------------------------------------------------------------------------------
    def _import_heightmap(self, data):
        import ctypes as C
        N, jitter = int(data["N"]), float(data["jitter"])
        if data.get("seed") is None:
            raise ValueError("importHeightmap needs a seed (the reference would draw Math.random())")
        seed, skip = data["seed"], bool(data.get("skipClimate"))
        gray = np.ascontiguousarray(data["grayscale"], np.uint8).reshape(-1)
        width, height = int(data["imageWidth"]), int(data["imageHeight"])
        if gray.size != width * height:
            raise ValueError("grayscale does not hold imageWidth * imageHeight pixels")
        temperatureOffset = data.get("temperatureOffset", 0) or 0
        precipitationOffset = data.get("precipitationOffset", 0) or 0
        landCoverage = data.get("landCoverage", 0.3) if data.get("landCoverage") is not None else 0.3
        timing = []

        def stage(name, t0):
            timing.append({"stage": name, "ms": 1e3 * (time.perf_counter() - t0)})

        t0 = time.perf_counter()
        mesh = self._building = DeviceMesh.build_sphere(N, jitter, seed, device=self.device, lib=self.lib, order=self.mesh_order)
        stage("Sphere mesh", t0); t0 = time.perf_counter()
        neighborDist = mesh.computeNeighborDist()
        stage("Neighbor distances", t0); t0 = time.perf_counter()
        t_xyz = mesh.generateTriangleCenters()
        triangles, halfedges = mesh.trianglesAndHalfedges()
        stage("Triangle centers", t0); t0 = time.perf_counter()
        n, dll = mesh.numRegions, mesh.lib.dll
        r_elevation = np.empty(n, np.float32)
        mesh._begin(r_elevation)
        mesh.lib.check(dll.pb_sample_heightmap(mesh._mesh, gray.ctypes.data, width, height, r_elevation.ctypes.data))
        stage("Sample heightmap", t0); t0 = time.perf_counter()
        prePostElev = r_elevation.copy()
        post = runPostProcessing(mesh, mesh.r_xyz, r_elevation, {k: data.get(k, 0) or 0 for k in _SLIDERS}, neighborDist, seed)
        stage("Terrain post-processing", t0); t0 = time.perf_counter()
        r_plate = np.empty(n, np.int32)
        mesh.lib.check(dll.pb_derive_synthetic_plates(mesh._mesh, r_elevation.ctypes.data, r_plate.ctypes.data))
        plateSeeds = [int(r) for r in np.nonzero(r_plate == np.arange(n))[0]]
        plateIsOcean = {s for s in plateSeeds if r_elevation[s] <= 0}
        plateVec = {s: [0, 0, 0] for s in plateSeeds}
        stage("Synthetic plates", t0)
        masks = [np.empty(n, np.uint8) for _ in range(3)]
        mesh.lib.check(dll.pb_classify_imported_regions(mesh._mesh, r_elevation.ctypes.data, *(m.ctypes.data for m in masks)))
        debugLayers = {"erosionDelta": post["dl_erosionDelta"]}
        wind = ocean = precip = temp = None
        if not skip:
            t0 = time.perf_counter()
            wind, ocean, precip, temp, _, _ = self._climate(mesh, r_elevation, plateIsOcean, r_plate, seed, temperatureOffset,
                                                            precipitationOffset, landCoverage, debugLayers)
            stage("Climate (wind, ocean currents, precipitation, temperature, Köppen)", t0)
        t0 = time.perf_counter()
        t_elevation = mesh.computeTriangleElevations(r_elevation)
        stage("Triangle elevations", t0)
        r_stress = np.zeros(n, np.float32)
        self._retain(dict(mesh=mesh, neighborDist=neighborDist, r_plate=r_plate.copy(), plateSeeds=list(plateSeeds), plateVec=plateVec,
                      plateIsOcean=set(plateIsOcean), originalPlateIsOcean=set(plateIsOcean), plateDensity={}, plateDensityLand={},
                      plateDensityOcean={}, prePostElev=prePostElev.copy(), r_elevation_final=r_elevation.copy(), seed=seed, nMag=0,
                      cachedWind=wind, cachedOcean=ocean))
        reply = {"type": "done", "triangles": triangles, "halfedges": halfedges, "numRegions": n, "r_xyz": mesh.r_xyz, "t_xyz": t_xyz,
                 "r_plate": r_plate, "plateSeeds": list(plateSeeds), "plateVec": plateVec,
                 "plateIsOcean": [s for s in plateSeeds if s in plateIsOcean], "originalPlateIsOcean": [s for s in plateSeeds if s in plateIsOcean],
                 "plateDensity": {}, "plateDensityLand": {}, "plateDensityOcean": {}, "prePostElev": prePostElev,
                 "r_elevation": r_elevation, "t_elevation": t_elevation, "mountain_r": _mask_to_list(masks[0]),
                 "coastline_r": _mask_to_list(masks[1]), "ocean_r": _mask_to_list(masks[2]), "r_stress": r_stress}
        reply.update(self._climate_fields(wind, ocean, precip, temp))
        reply.update(skipClimate=skip, seed=seed, nMag=0, debugLayers=debugLayers, _timing=[], _pipelineTiming=timing,
                     _postTiming=post.get("postTiming", []), _workerTotal=sum(s["ms"] for s in timing),
                     _params={"N": N, "P": 0, "jitter": jitter, "nMag": 0, "numContinents": 0, "seed": seed,
                              **{k: data.get(k) for k in _SLIDERS}})
        return reply

    # ---- reapply (:341-440) ----------------------------------------------------------------------------------------------
    def _reapply(self, data):
        W, skip = self.W, bool(data.get("skipClimate"))
        temperatureOffset, precipitationOffset, landCoverage = self._climate_params(data)
        mesh = W["mesh"]
        t0 = time.perf_counter()
        r_elevation = W["prePostElev"].copy()
        post = runPostProcessing(mesh, mesh.r_xyz, r_elevation, {k: data.get(k, 0) or 0 for k in _SLIDERS}, W["neighborDist"], W["seed"])
        tPost = 1e3 * (time.perf_counter() - t0)
        W["r_elevation_final"] = r_elevation.copy()
        wind = ocean = precip = temp = None
        layers, t = None, {}
        if not skip:
            layers = {}
            wind, ocean, precip, temp, _, t = self._climate(mesh, r_elevation, W["plateIsOcean"], W["r_plate"], W["seed"], temperatureOffset,
                                                            precipitationOffset, landCoverage, layers)
            layers.pop("continentality")               # windDebugLayers has no continentality entry (:409-421)
        W["cachedWind"], W["cachedOcean"] = wind, ocean
        reply = {"type": "reapplyDone", "skipClimate": skip, "r_elevation": r_elevation,
                 "t_elevation": mesh.computeTriangleElevations(r_elevation), "erosionDelta": post["dl_erosionDelta"]}
        reply.update(self._climate_fields(wind, ocean, precip, temp))
        reply.update(windDebugLayers=layers, _reapplyTiming=dict(t, postProcessing=tPost), _postTiming=post.get("postTiming", []))
        return reply

    # ---- editRecompute (:442-577) ------------------------------------------------------------------------------------------
    def _edit_recompute(self, data):
        W, skip = self.W, bool(data.get("skipClimate"))
        temperatureOffset, precipitationOffset, landCoverage = self._climate_params(data)
        W["plateIsOcean"] = set(data["plateIsOcean"])
        W["plateDensity"] = {int(k): float(v) for k, v in data["plateDensity"].items()}
        mesh, pio, r_plate, seeds, vec, seed = W["mesh"], W["plateIsOcean"], W["r_plate"], W["plateSeeds"], W["plateVec"], W["seed"]
        superPlateData = None
        if (W.get("P") or 0) >= 8:
            superPlateData = pl.buildSuperPlates(mesh, r_plate, seeds, vec, pio, W["plateDensity"])
        t0 = time.perf_counter()
        res = assignElevation(mesh, mesh.r_xyz, pio, r_plate, vec, seeds, seed, float(data["nMag"]), seed, 5, W["plateDensity"], superPlateData)
        tElev = 1e3 * (time.perf_counter() - t0)
        r_elevation, debugLayers = res["r_elevation"], res["debugLayers"]
        if superPlateData is not None:
            debugLayers["superPlates"] = np.asarray(superPlateData["r_superPlate"], np.float32)
        prePostElev = r_elevation.copy()
        t0 = time.perf_counter()
        post = runPostProcessing(mesh, mesh.r_xyz, r_elevation, {k: data.get(k, 0) or 0 for k in _SLIDERS}, W["neighborDist"], seed,
                                 debugLayers["hotspot"])
        tPost = 1e3 * (time.perf_counter() - t0)
        debugLayers["erosionDelta"] = post["dl_erosionDelta"]
        W["r_elevation_final"] = r_elevation.copy()
        wind = ocean = precip = temp = None
        t = {}
        if not skip:
            wind, ocean, precip, temp, _, t = self._climate(mesh, r_elevation, pio, r_plate, seed, temperatureOffset, precipitationOffset,
                                                            landCoverage, debugLayers)
        W["cachedWind"], W["cachedOcean"] = wind, ocean
        W["prePostElev"] = prePostElev.copy()
        reply = {"type": "editDone", "skipClimate": skip, "prePostElev": prePostElev, "r_elevation": r_elevation,
                 "t_elevation": mesh.computeTriangleElevations(r_elevation),
                 "mountain_r": _mask_to_list(res["mountain_r"]), "coastline_r": _mask_to_list(res["coastline_r"]),
                 "ocean_r": _mask_to_list(res["ocean_r"]), "r_stress": res["r_stress"]}
        reply.update(self._climate_fields(wind, ocean, precip, temp))
        reply.update(debugLayers=debugLayers, _editTiming=dict(t, elevation=tElev, postProcessing=tPost), _timing=res.get("_timing", []),
                     _postTiming=post.get("postTiming", []))
        return reply

    # ---- computeClimate (:579-677) -------------------------------------------------------------------------------------------
    def _compute_climate(self, data):
        W = self.W
        temperatureOffset, precipitationOffset, landCoverage = self._climate_params(data)
        mesh, elev = W["mesh"], W["r_elevation_final"]
        wind, ocean = W["cachedWind"], W["cachedOcean"]
        t = {"wind": 0.0, "ocean": 0.0}
        if wind is None:
            t0 = time.perf_counter()
            wind = cl.computeWind(mesh, mesh.r_xyz, elev, W["plateIsOcean"], W["r_plate"], W["seed"])
            t["wind"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
            ocean = cl.computeOceanCurrents(mesh, mesh.r_xyz, elev, wind)
            t["ocean"] = 1e3 * (time.perf_counter() - t0)
            W["cachedWind"], W["cachedOcean"] = wind, ocean
        t0 = time.perf_counter()
        precip = cl.computePrecipitation(mesh, mesh.r_xyz, elev, wind, ocean, precipitationOffset, landCoverage)
        t["precipitation"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        temp = cl.computeTemperature(mesh, mesh.r_xyz, elev, wind, ocean, precip, temperatureOffset)
        t["temperature"] = 1e3 * (time.perf_counter() - t0); t0 = time.perf_counter()
        koppen = cl.classifyKoppen(mesh, elev, temp, precip)
        t["koppen"] = 1e3 * (time.perf_counter() - t0)
        reply = {"type": "climateDone"}
        reply.update(self._climate_fields(wind, ocean, precip, temp))
        reply["climateDebugLayers"] = dict(
            pressureSummer=wind["r_pressure_summer"], pressureWinter=wind["r_pressure_winter"],
            windSpeedSummer=wind["r_wind_speed_summer"], windSpeedWinter=wind["r_wind_speed_winter"],
            continentality=wind["r_continentality"], precipSummer=precip["r_precip_summer"], precipWinter=precip["r_precip_winter"],
            rainShadowSummer=precip["r_rainshadow_summer"], rainShadowWinter=precip["r_rainshadow_winter"],
            tempSummer=temp["r_temperature_summer"], tempWinter=temp["r_temperature_winter"], koppen=koppen)
        reply["_climateTiming"] = t
        return reply
