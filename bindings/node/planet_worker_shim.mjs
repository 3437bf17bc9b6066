// planet_worker_shim.mjs — the reference's stage functions (same names, same signatures, same result keys)
// implemented over the N-API addon.  js/planet-worker.js keeps its handlers unchanged and only swaps its imports:
//
//   - import { warpTerrain, smoothElevation, erodeComposite, sharpenRidges, applySoilCreep } from './terrain-post.js';
//   - import { assignElevation } from './elevation.js';
//   - import { computeWind } from './wind.js';  import { computeOceanCurrents } from './ocean.js';  …
//   + import { warpTerrain, …, assignElevation, computeWind, … } from '../bindings/node/planet_worker_shim.mjs';
//
// Every name js/planet-worker.js imports (:4-17) is exported here, so the swap is the import lines and nothing else: makeRng and
// SimplexNoise come from here too (they carry the seed the library needs), buildSphere returns a mesh object with the SphereMesh
// members the handlers read.  (Node worker_threads build of the worker; the browser build keeps the JS modules.)  No Node exists
// in this image: tests/test_zz_dropin_worker.py executes the unmodified reference worker over THIS file with the import lines
// redirected, under the evaluator tests/golden/minijs.py and the in-process Node-API runtime tests/napi_host/.
import { createRequire } from 'node:module';
const native = createRequire(import.meta.url)('./build/Release/planet_b200_addon.node');

export function setMesh(mesh, r_xyz) { native.setMesh(mesh.numRegions, mesh.adjOffset, mesh.adjList, r_xyz); }
export const setOption = native.setOption;
export const computeNeighborDist = native.computeNeighborDist;      // (mesh, r_xyz) → Float32Array
export const warpTerrain = native.warpTerrain;                      // in place, like the reference
export const smoothElevation = native.smoothElevation;
export const erodeComposite = native.erodeComposite;
export const sharpenRidges = native.sharpenRidges;
export const applySoilCreep = native.applySoilCreep;
export const smoothField = native.smoothField;

// What the worker builds itself and hands to the stage functions only as a carrier of its seed (js/planet-worker.js:146, 203):
// makeRng(seed) and new SimplexNoise(seed).  Imported from here they remember the seed; the streams themselves run in the library.
export function makeRng(seed) {          // js/rng.js:3-6 (Park–Miller on a hashed seed); the worker itself draws the plate densities from it (:196-199)
  let state = Math.abs(Math.floor(seed * 9301 + 49297)) % 2147483646 + 1;
  const rng = () => ((state = state * 16807 % 2147483647) - 1) / 2147483646;
  rng.seed = seed;
  return rng;
}
export class SimplexNoise { constructor(seed) { this.seed = seed; } }
export function setDelaunator() {}        // the triangulator is the library's (option mesh_order)

// The members of SphereMesh (js/sphere-mesh.js:94-172) that the worker and the main thread read.
class RetainedMesh {
  constructor(r) {
    const t = native.getMeshTriangles();
    this.numRegions = r.numRegions; this.adjOffset = this._adjOffset = r.adjOffset; this.adjList = this._adjList = r.adjList;
    this.triangles = t.triangles; this.halfedges = t.halfedges;
    this.numSides = t.triangles.length; this.numTriangles = (t.triangles.length / 3) | 0;
  }
  _next(s) { return (s % 3 === 2) ? s - 2 : s + 1; }
  s_begin_r(s) { return this.triangles[s]; }
  s_end_r(s) { return this.triangles[this._next(s)]; }
  s_inner_t(s) { return (s / 3) | 0; }
  s_outer_t(s) { return (this.halfedges[s] / 3) | 0; }
  r_circulate_r(out, r) {
    const start = this.adjOffset[r], len = this.adjOffset[r + 1] - start;
    out.length = len;
    for (let i = 0; i < len; i++) out[i] = this.adjList[start + i];
    return out;
  }
}
export { RetainedMesh as SphereMesh };

// buildSphere(N, jitter, rng) with rng = makeRng(seed) from this module.  Returns {mesh, r_xyz} like js/sphere-mesh.js:174; the
// mesh is also retained for the stage functions below.
export function buildSphere(N, jitter, rng) {
  const r = native.buildSphereFlat(N, jitter, rng.seed);
  return { mesh: new RetainedMesh(r), r_xyz: r.r_xyz };
}
export function generateTriangleCenters(mesh, r_xyz) { return native.generateTriangleCenters(); }

// generateCoarsePlates(seed, numPlates, numContinents, continentSizeVariety, landCoverage)      js/coarse-plates.js:19
export function generateCoarsePlates(seed, numPlates, numContinents, continentSizeVariety = 0, landCoverage = 0.3) {
  const r = native.generateCoarsePlatesFlat(seed, numPlates, numContinents, continentSizeVariety, landCoverage);
  const seeds = Array.from(r.seeds.subarray(0, r.numPlates));
  const coarsePlateVec = {}, coarsePlateIsOcean = new Set();
  seeds.forEach((pid, k) => {
    coarsePlateVec[pid] = { pole: Array.from(r.pole.subarray(3 * k, 3 * k + 3)), omega: r.omega[k] };
    if (r.isOcean[k]) coarsePlateIsOcean.add(pid);
  });
  return { coarseMesh: { numRegions: r.numRegions, adjOffset: r.adjOffset, adjList: r.adjList }, coarse_xyz: r.coarse_xyz,
           coarse_r_plate: r.coarse_r_plate, coarsePlateSeeds: new Set(seeds), coarsePlateVec, coarsePlateIsOcean };
}
export function projectCoarsePlates(mesh, r_xyz, coarseMesh, coarse_xyz, coarse_r_plate, seed, numPlates) {
  return native.projectCoarsePlatesFlat(coarseMesh.numRegions, coarseMesh.adjOffset, coarseMesh.adjList, coarse_xyz, coarse_r_plate,
                                        seed, numPlates ?? null);
}
export function smoothAndReconnectPlates(mesh, r_plate, plateSeeds, numPasses) {
  native.smoothAndReconnectPlatesFlat(r_plate, Int32Array.from(plateSeeds), numPasses);
}
export function buildSuperPlates(mesh, r_plate, plateSeeds, plateVec, plateIsOcean, plateDensity) {
  const ids = Int32Array.from(plateSeeds), n = ids.length;
  const isOcean = new Uint8Array(n), pole = new Float64Array(3 * n).fill(NaN), omega = new Float64Array(n), dens = new Float64Array(n).fill(NaN);
  ids.forEach((pid, k) => {
    isOcean[k] = plateIsOcean.has(pid) ? 1 : 0;
    const pv = plateVec[pid];
    if (pv && pv.pole) { pole.set(pv.pole, 3 * k); omega[k] = pv.omega; }
    if (plateDensity[pid] !== undefined) dens[k] = plateDensity[pid];
  });
  const r = native.buildSuperPlatesFlat(r_plate, ids, isOcean, pole, omega, dens);
  const superPlateVec = {}, superPlateIsOcean = new Set(), superPlateDensity = {};
  for (let sp = 0; sp < r.numSuperPlates; sp++) {
    superPlateVec[sp] = { pole: Array.from(r.pole.subarray(3 * sp, 3 * sp + 3)), omega: r.omega[sp] };
    if (r.isOcean[sp]) superPlateIsOcean.add(sp);
    superPlateDensity[sp] = r.density[sp];
  }
  return { r_superPlate: r.r_superPlate, superPlateVec, superPlateIsOcean, superPlateDensity, numSuperPlates: r.numSuperPlates };
}

// runPostProcessing is module-private in planet-worker.js (:40-102); the worker may call this instead.
export function runPostProcessing(mesh, r_xyz, r_elevation, params, neighborDist, seed, r_hotspot) {
  const r = native.runPostProcessing(mesh, r_xyz, r_elevation, params, neighborDist, seed, r_hotspot ?? null);
  return { dl_erosionDelta: r.dl_erosionDelta,
           postTiming: Object.entries(r.postTimingMs ?? {}).map(([stage, ms]) => ({ stage, ms })) };
}

function flattenPlates(isOceanSet, vec, density, order) {
  const ids = Int32Array.from(order ?? Object.keys(vec).map(Number));
  const n = ids.length;
  const isOcean = new Uint8Array(n), pole = new Float64Array(3 * n), omega = new Float64Array(n), dens = new Float64Array(n);
  ids.forEach((pid, k) => {
    isOcean[k] = isOceanSet.has(pid) ? 1 : 0;
    pole.set(vec[pid].pole, 3 * k); omega[k] = vec[pid].omega; dens[k] = density[pid];
  });
  return [ids, isOcean, pole, omega, dens];
}

// assignElevation(mesh, r_xyz, plateIsOcean, r_plate, plateVec, plateSeeds, noise, noiseMag, seed, spread, plateDensity, superPlateData)
// `noise` must expose the seed it was built from (new SimplexNoise(seed), js/planet-worker.js:203): pass {seed}.
export function assignElevation(mesh, r_xyz, plateIsOcean, r_plate, plateVec, plateSeeds, noise, noiseMag, seed, spread,
                                plateDensity, superPlateData) {
  const args = [r_plate, ...flattenPlates(plateIsOcean, plateVec, plateDensity), Int32Array.from(plateSeeds),
                noise.seed, noiseMag, seed, spread];
  if (superPlateData) {
    args.push(superPlateData.r_superPlate,
              ...flattenPlates(superPlateData.superPlateIsOcean, superPlateData.superPlateVec, superPlateData.superPlateDensity));
  }
  const r = native.assignElevationFlat(...args);
  const toSet = (mask) => { const s = new Set(); for (let i = 0; i < mask.length; i++) if (mask[i]) s.add(i); return s; };
  if (superPlateData) r.debugLayers.superPlates = new Float32Array(superPlateData.r_superPlate);
  // the reference returns insertion-ordered Sets; downstream code only tests membership / draws the regions
  return { r_elevation: r.r_elevation, mountain_r: toSet(r.mountain_r), coastline_r: toSet(r.coastline_r),
           ocean_r: toSet(r.ocean_r), r_stress: r.r_stress, debugLayers: r.debugLayers, _timing: [] };
}

// result objects fetch their arrays from the device on first access
function lazyResult(keys, extra = {}) {
  const cache = { ...extra };
  return new Proxy(cache, { get: (t, k) => (k in t ? t[k] : (keys.includes(k) ? (t[k] = native.getClimateField(k)) : undefined)),
                            has: (t, k) => k in t || keys.includes(k), ownKeys: () => [...new Set([...Object.keys(cache), ...keys])] });
}
const S = ['summer', 'winter'];
const WIND_KEYS = [...S.flatMap(s => [`r_pressure_${s}`, `r_wind_east_${s}`, `r_wind_north_${s}`, `r_wind_speed_${s}`]),
  'itczLons', 'itczLatsSummer', 'itczLatsWinter', 'r_lat', 'r_lon', 'r_sinLat', 'r_isLand', 'r_continentality', 'r_coastDistLand',
  'r_plateContinentality', 'r_eastX', 'r_eastY', 'r_eastZ', 'r_northX', 'r_northY', 'r_northZ'];
const OCEAN_KEYS = S.flatMap(s => [`r_ocean_current_east_${s}`, `r_ocean_current_north_${s}`, `r_ocean_speed_${s}`, `r_ocean_warmth_${s}`]);

export function computeWind(mesh, r_xyz, r_elevation, plateIsOcean, r_plate, noise, axialTilt = 23.5) {
  native.computeWindFlat(r_elevation, Int32Array.from(plateIsOcean), r_plate, noise.seed, axialTilt);
  return lazyResult(WIND_KEYS, { _windTiming: [] });
}
export function computeOceanCurrents(mesh, r_xyz, r_elevation, windResult) {
  native.computeOceanCurrentsFlat(r_elevation);
  return lazyResult(OCEAN_KEYS, { _oceanTiming: [] });
}
export function computePrecipitation(mesh, r_xyz, r_elevation, windResult, oceanResult, precipitationOffset = 0, landCoverage = 0.3) {
  native.computePrecipitationFlat(r_elevation, precipitationOffset, landCoverage);
  return lazyResult(S.flatMap(s => [`r_precip_${s}`, `r_rainshadow_${s}`]), { _precipTiming: [] });
}
export function computeTemperature(mesh, r_xyz, r_elevation, windResult, oceanResult, precipResult, temperatureOffset = 0) {
  native.computeTemperatureFlat(r_elevation, temperatureOffset);
  return lazyResult(S.map(s => `r_temperature_${s}`), { _tempTiming: [] });
}
export function classifyKoppen(mesh, r_elevation, tempResult, precipResult) { return native.classifyKoppenFlat(r_elevation); }

// exportMap's pixels (js/planet-mesh.js:1752-1950) for the main thread: `ctx.putImageData(new ImageData(px, width), 0, 0)` and
// `canvas.toBlob` replace the tiled WebGL render + readback; type is the reference's export type name.
export function exportMapPixels(type, width, r_elevation, r_koppen = null) { return native.exportMapPixels(type, width, r_elevation, r_koppen); }
