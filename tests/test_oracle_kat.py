"""Pins the oracle against the derived known-answer vectors of SURVEY.md §8(c).  The reference ships
no tests or golden vectors and no JS runtime exists here; these vectors were derived by hand from the cited lines.  (Since
round 2 the reference's own rng.js / simplex-noise.js reproduce them under tests/golden/minijs.py, and whole worker replies
are pinned in tests/test_zz_reference_vectors.py.)"""
import numpy as np


def test_rng(oracle):
    np.testing.assert_array_equal(oracle.rng(0, 4), [0.3858243514651659, 0.5498798815998062, 0.8311735706694178, 0.5342035619860548])
    np.testing.assert_array_equal(oracle.rng(42, 4), [0.4431328917323918, 0.7345157044329826, 0.005446482454842406, 0.5390384020647392])
    np.testing.assert_array_equal(oracle.rng(42.5, 4), [0.4795255088056675, 0.38523056999336047, 0.5701946896223302, 0.2621518459749891])


def test_simplex(oracle):
    perm, pm12 = oracle.simplex_perm(42)
    assert perm[:12].tolist() == [124, 100, 59, 193, 92, 16, 78, 212, 47, 194, 101, 93]
    assert (pm12 == perm % 12).all() and (perm[256:] == perm[:256]).all()
    p = [[0.1, 0.2, 0.3]]
    assert oracle.noise(42, "noise3D", p)[0] == -0.11666551466666661
    assert oracle.noise(42, "fbm", p)[0] == -0.006157208265402842
    assert oracle.noise(42, "fbm", p, octaves=3, persistence=0.5)[0] == 0.05088801219047616
    assert oracle.noise(42, "ridgedFbm", p, octaves=6)[0] == 0.535666463298141
    assert oracle.noise(42, "noise3D", [[-1.7, 2.4, 0.05]])[0] == 0.37858406795833355
    assert oracle.noise(42, "noise3D", [[4, 4, 4]])[0] == 0
    assert oracle.noise(42, "ridgedFbm", [[4, 4, 4]], octaves=6)[0] == 1


def test_cell_noise_js_uint32_semantics(oracle):
    assert oracle.cell_noise(1) == 0.0038813726519889603
    assert oracle.cell_noise(1000) == 0.002694108784826032
    assert oracle.cell_noise(5000000) == 0.004767057684894432
    assert oracle.cell_noise(49999999) == 0.009582771253674937   # a uint32 wrap-around port gives 0.00316…


def test_detmath_close_to_libm(oracle):
    rng = np.random.default_rng(1)
    x = rng.uniform(1e-6, 50, 20000)
    y = rng.uniform(0.1, 2.0, 20000)
    for kind, ref in (("exp", np.exp(x / 10)), ("log", np.log(x)), ("pow", np.power(x, y)), ("sin", np.sin(x)),
                      ("cos", np.cos(x)), ("atan", np.arctan(x)), ("tanh", np.tanh(x / 10))):
        arg = x / 10 if kind in ("exp", "tanh") else x
        got = oracle.detmath(kind, arg, y)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        assert rel.max() < 1e-14, (kind, rel.max())
    u = rng.uniform(-1, 1, 20000)
    assert np.abs(oracle.detmath("asin", u) - np.arcsin(u)).max() < 1e-15
    assert (oracle.detmath("pow", x, np.full_like(x, 0.5)) == np.sqrt(x)).all()


def test_percentile_and_smooth_field(oracle, planet_small):
    mesh, xyz, nd, elev = planet_small()
    f = elev.copy()
    oracle.smooth_field(mesh, f, 3)
    # numpy restatement of js/climate-util.js:5-25
    g = elev.copy()
    deg = np.diff(mesh.adjOffset)
    rows = np.repeat(np.arange(mesh.numRegions), deg)
    for _ in range(3):
        s = g.astype(np.float64)
        # sequential left-to-right accumulation per row, as the reference does
        acc = s.copy()
        maxd = deg.max()
        for k in range(maxd):
            has = deg > k
            idx = mesh.adjOffset[:-1][has] + k
            acc[has] += s[mesh.adjList[idx]]
        g = (acc / (deg + 1)).astype(np.float32)
    np.testing.assert_array_equal(f, g)
    assert oracle.percentile(elev, 0.95) == np.sort(elev)[int(np.floor(elev.size * 0.95))]
