"""CPU tests of the CT-number front end: monte_hu_classes_default / monte_ctnum_segment (host helpers of
libmonte_gpu, no device needed) and the present-material majorant against the oracle's.  The transport on a segmented
HU volume is checked history by history in tests/test_emu_mc.py (emulated kernel) and tests/test_mc_gpu.py (B200)."""
import ctypes as C

import numpy as np
import pytest

from monte_b200 import _abi, api, scenes


def test_default_classes_and_segmentation_against_numpy():
    classes = api.hu_classes_default(True)
    assert len(classes) == 8
    edges = np.array([c.hu_min for c in classes], np.float32)
    assert np.all(np.diff(edges) > 0) and edges[0] == -900
    rng = np.random.default_rng(3)
    hu = rng.uniform(-1100, 2500, size=(7, 9, 11)).astype(np.float32)
    hu.flat[:8] = edges                                   # a value on an edge belongs to the class that starts there
    hu.flat[8] = np.nan                                   # NaN: air
    base = scenes.make_xs()
    lab, xs, mu, present = api.ctnum_segment(hu, classes, base, 80.0)
    want = np.searchsorted(edges, hu, side="right").astype(np.uint8)
    want[np.isnan(hu)] = 0
    assert np.array_equal(lab, want)
    assert present == sum(1 << (k - 1) for k in np.unique(want) if k)
    assert xs.n_materials == 8
    # tables: mixture rule on the mass coefficients, the class's own density
    for m, c in enumerate(classes):
        f = c.frac_b if c.material_b >= 0 else 0.0
        for name in ("coh", "compt", "photo", "total"):
            a = np.ctypeslib.as_array(getattr(base, name))
            got = np.ctypeslib.as_array(getattr(xs, name))[m]
            mix = (1.0 - f) * a[c.material_a].astype(np.float64) + (f * a[c.material_b].astype(np.float64) if c.material_b >= 0 else 0.0)
            assert np.array_equal(got, mix.astype(np.float32)), (m, name)
        assert xs.density[m] == np.float32(c.density)
    # mu is what the transport sees at that energy
    tot = np.ctypeslib.as_array(xs.total)
    mu_of = np.concatenate([[0.0], [np.float32(tot[m][80]) * np.float32(xs.density[m]) for m in range(8)]]).astype(np.float32)
    assert np.array_equal(mu, mu_of[lab])
    # water at HU 0: mu = mu_water * 1.01 (the soft-tissue bin), bone classes are denser AND carry calcium
    assert lab[np.unravel_index(np.nanargmin(np.abs(hu)), hu.shape)] == 3
    assert tot[7][60] > tot[4][60] > tot[2][60]
    # without a calcium table the bone classes are water at their density
    lab2, xs2, _, _ = api.ctnum_segment(hu, api.hu_classes_default(False), scenes.make_xs(("h2o",)), 80.0)
    assert np.array_equal(lab2, lab)
    assert np.array_equal(np.ctypeslib.as_array(xs2.total)[7], np.ctypeslib.as_array(base.total)[0])


def test_segmentation_classes_with_air_gaps_and_errors():
    HuClass = _abi.HuClass
    base = scenes.make_xs()
    # air | water | an air class in the middle (density 0) | calcium-loaded
    cl = (HuClass * 3)(HuClass(-500, 0, -1, 0, 1.0), HuClass(100, 0, -1, 0, 0.0), HuClass(300, 0, 1, 0.5, 1.5))
    hu = np.array([-800, -500, 0, 99, 100, 250, 300, 5000], np.float32)
    lab, xs, mu, present = api.ctnum_segment(hu, cl, base, 60.0)
    assert lab.tolist() == [0, 1, 1, 1, 0, 0, 2, 2] and xs.n_materials == 2 and present == 0b11
    assert mu[4] == 0 and mu[6] == np.float32(xs.total[1][60]) * np.float32(1.5)
    bad = (HuClass * 2)(HuClass(0, 0, -1, 0, 1.0), HuClass(0, 0, -1, 0, 1.0))
    with pytest.raises(api.MonteError, match="ascend"):
        api.ctnum_segment(hu, bad, base, 60.0)
    with pytest.raises(api.MonteError, match="outside the base tables"):
        api.ctnum_segment(hu, (HuClass * 1)(HuClass(0, 5, -1, 0, 1.0)), base, 60.0)
    with pytest.raises(api.MonteError, match="air"):
        api.ctnum_segment(hu, (HuClass * 1)(HuClass(0, 0, -1, 0, 0.0)), base, 60.0)
    many = (HuClass * 9)(*[HuClass(100.0 * i, 0, -1, 0, 1.0 + 0.1 * i) for i in range(9)])
    with pytest.raises(api.MonteError, match="non-air classes"):
        api.ctnum_segment(hu, many, base, 60.0)


def test_present_material_majorant_matches_monte_xs_majorant(oracle):
    """the majorant the oracle tracks with under MONTE_MC_MAJORANT_PRESENT is monte_xs_majorant of the labels: fewer
    tentative collisions than with the majorant of all tables, same expected tallies (MT19937 runs, chi-square)"""
    keV = 60.0
    hu = scenes.hu_head_phantom(33, 0.6)
    lab, xs, mu, present = api.ctnum_segment(hu, api.hu_classes_default(True), scenes.make_xs(), keV)
    mm = np.zeros(_abi.TABLE_ROWS, np.float32)
    assert api.load().monte_xs_majorant(C.byref(xs), lab.ctypes.data, lab.size, mm.ctypes.data) == 0
    tot = np.ctypeslib.as_array(xs.total)
    want = max(np.float32(tot[m][60]) * np.float32(xs.density[m]) for m in range(8) if present >> m & 1)
    assert mm[60] == want and mm[60] < np.float32(tot[7][60]) * np.float32(xs.density[7])
    g = scenes.mc_geom(17, 32.5 / 17, n_views=1)
    vol = scenes.volume_for(lab, 0.6)
    spec = scenes.mono_spectrum(keV)
    per = 400
    res = {}
    for mode in (_abi.MAJORANT_ALL, _abi.MAJORANT_PRESENT):
        vol.majorant_mode = mode
        res[mode] = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_MT, seed=5 + mode, n_threads=4), per)
    a0, a5, ra = res[_abi.MAJORANT_ALL][:3]
    p0, p5, rp = res[_abi.MAJORANT_PRESENT][:3]
    ratio = rp["woodcock_steps"] / ra["woodcock_steps"]
    assert ratio < 0.9, ratio
    for k in ("primaries", "absorbed", "interactions", "scatter_detected"):
        assert abs(ra[k] - rp[k]) <= 5.0 * np.sqrt(ra[k] + rp[k] + 1), (k, ra[k], rp[k])
    d = (a0.astype(np.float64) - p0), (a0 + p0).astype(np.float64)
    m = d[1] > 0
    var = d[1][m] * (1.0 - d[1][m] / (2.0 * per))
    c2 = (d[0][m] ** 2 / np.maximum(var, 1e-9)).sum()
    assert c2 < m.sum() + 6.0 * np.sqrt(2.0 * m.sum()), (c2, m.sum())
    # an all-air volume under PRESENT: transparent, every photon is a primary (and the oracle returns)
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    air = np.zeros_like(lab)
    e0, e5, re = oracle.mc_run(g, vol, air, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_PHILOX, seed=1), 3)[:3]
    assert re["primaries"] == re["histories"] == 17 * 17 * 3 and re["interactions"] == 0
