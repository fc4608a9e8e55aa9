"""CPU: the Monte-Carlo oracle against the reference's own outputs and analytic known answers.

The strongest pin: with every quirk of the shipped monte_cpp/CBCT_real2.cpp switched on and the
reference's MT19937 seeded as the binary is (oracle/shim/windows.h pins time() to 5489), the oracle
reproduces the unmodified binary's 65x65 scatter image and its three printed counters bit for bit.
The golden (tests/golden/mc_real2.npz) was produced by tests/golden/make_golden_mc.py.
"""
import math
import os

import numpy as np
import pytest

from monte_b200 import _abi, scenes
from conftest import GOLDEN


def real2_scene():
    """the as-shipped run: spher01.raw geometry (make_image01.cpp), lookup box of CBCT_real2.cpp:770"""
    sphere = np.zeros((325, 185, 185), np.uint8)
    kk, jj, ii = np.ogrid[:325, :185, :185]
    sphere[((ii - 90) ** 2 + (jj - 90) ** 2 + (kk - 160) ** 2) <= 2500] = 1      # make_image01.cpp:19
    vol = _abi.McVolume()
    vol.nx, vol.ny, vol.nz, vol.pitch = 185, 185, 325, 0.1
    vol.origin[0] = vol.origin[1] = -9.05
    vol.origin[2] = -16.05
    for a, (lo, hi) in enumerate(((-6, 10), (-6, 6), (-10, 10))):
        vol.clip_lo[a], vol.clip_hi[a] = lo, hi
    return scenes.mc_geom(65, 0.5), vol, sphere


def test_philox2x32_known_answers(oracle):
    """Random123 kat_vectors, philox2x32 10 rounds"""
    assert oracle.philox2x32(0, 0, 0) == (0xff1dae59, 0x6cd10df2)
    assert oracle.philox2x32(0xffffffff, 0xffffffff, 0xffffffff) == (0x2c3f628b, 0xab4fd7ad)
    assert oracle.philox2x32(0x243f6a88, 0x85a308d3, 0x13198a2e) == (0xdd7ce038, 0xf62a4c12)


def test_table_known_answers():
    """SURVEY 8a-A1"""
    h2o, ca = scenes.load_tables()
    assert abs(h2o[3, 140] - 0.1538092) < 1e-9 and abs(ca[3, 140] - 0.17730) < 1e-9
    assert abs(h2o[3, 60] - 0.20585) < 1e-5
    hq, cq = scenes.load_tables(quirk_bom=True)
    assert hq[0, 1] == 1.372 and cq[0, 1] == 1.372          # CBCT_real2.cpp:663


def test_oracle_reproduces_unmodified_cbct_real2_bit_for_bit(oracle):
    gold = np.load(os.path.join(GOLDEN, "mc_real2.npz"))
    g, vol, sphere = real2_scene()
    assert int(sphere.sum()) == int(gold["sphere_voxels"])
    h2o, ca = scenes.load_tables(quirk_bom=True)
    tb = oracle.tables_from_arrays([(h2o, 1.0), (ca, 1.55)])
    o = oracle.mc_opts(rng_mode=oracle.RNG_MT, quirks=oracle.Q_REAL2, seed=int(gold["seed"]), n_threads=1)
    im0, im5, res, _, _ = oracle.mc_run(g, vol, sphere, tb, scenes.mono_spectrum(140.0), o, int(gold["per"]),
                                        views=(0, 1), pixels=(32, 33, 32, 33))
    assert np.array_equal(im5[0], gold["image"])
    assert res["num_scatter"] == int(gold["num_scatter"])
    assert res["num_nd"] == int(gold["num_nd"])
    assert res["scatter_detected"] == int(gold["count"])
    assert not im0.any()                                   # CBCT_real2.cpp:290: primaries not tallied
    # coarse known answers of SURVEY 8c(4): 75.07 % and 1.13 %
    assert abs(res["num_scatter"] / 1e7 - 0.7507) < 5e-4


@pytest.mark.slow
def test_golden_is_what_the_binary_prints(oracle):
    if not oracle.have_ref("CBCT_real2"):
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    gold = np.load(os.path.join(GOLDEN, "mc_real2.npz"))
    img, c, _ = oracle.ref_cbct_real2()
    assert np.array_equal(img, gold["image"]) and c["count"] == int(gold["count"])


def small_scene(n=33, pitch=1.0, det=17, views=4, mode=_abi.SOURCE_PENCIL):
    lab = scenes.cylinder_phantom(n, pitch)
    g = scenes.mc_geom(det, 32.5 / det, n_views=views, source_mode=mode)
    g.angle_step_deg = 90.0 / views * 4
    return g, scenes.volume_for(lab, pitch), lab


def test_primary_transmission_known_answer(oracle):
    """exp(-mu*L): central pencil through 20 cm of water (SURVEY 8c(4) analytic KAT), both RNGs"""
    lab = scenes.cylinder_phantom(41, 0.5, rods=False)
    g = scenes.mc_geom(1, 0.5, n_views=1)
    g.half = 0.25
    vol = scenes.volume_for(lab, 0.5)
    xs = scenes.make_xs()
    tb = oracle.tables_from_xs(xs)
    per = 400000
    chord = 20.5                                             # 41 voxels of 0.5 cm across the diameter
    p = math.exp(-float(xs.total[0][140]) * chord)
    for mode in (oracle.RNG_MT, oracle.RNG_PHILOX):
        im0, im5, res, _, _ = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0),
                                            oracle.mc_opts(rng_mode=mode, seed=3), per)
        assert res["primaries"] == im0.sum()
        assert abs(res["primaries"] / per - p) < 4 * math.sqrt(p * (1 - p) / per)
        assert res["histories"] == per and im5.sum() == res["primaries"] + res["scatter_detected"]


def test_mt_and_philox_modes_agree_statistically(oracle):
    g, vol, lab = small_scene()
    tb = oracle.tables_from_xs(scenes.make_xs())
    per = 3000
    a = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), oracle.mc_opts(oracle.RNG_MT, seed=5), per)
    b = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), oracle.mc_opts(oracle.RNG_PHILOX, seed=5), per)
    for k in ("primaries", "scatter_detected", "absorbed", "compton", "coherent"):
        na, nb = a[2][k], b[2][k]
        assert abs(na - nb) < 5 * math.sqrt(na + nb + 1), k
    sa, sb = a[1].astype(np.int64) - a[0], b[1].astype(np.int64) - b[0]       # scatter-only images
    chi2 = ((sa - sb) ** 2 / np.maximum(sa + sb, 1))[(sa + sb) > 0]
    assert abs(chi2.sum() - chi2.size) < 5 * math.sqrt(2 * chi2.size)


def test_philox_result_is_partition_independent(oracle):
    g, vol, lab = small_scene(views=2)
    tb = oracle.tables_from_xs(scenes.make_xs())
    o = oracle.mc_opts(oracle.RNG_PHILOX, seed=9)
    whole = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(), o, 40)
    p1 = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(), o, 40, n_range=(0, 13))
    p2 = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(), o, 40, n_range=(13, 40))
    assert np.array_equal(whole[0], p1[0] + p2[0]) and np.array_equal(whole[1], p1[1] + p2[1])


def test_project_primary_chord_lengths(oracle):
    """line integral through the water cylinder = mu * chord (voxelisation error ~ one voxel)"""
    lab = scenes.cylinder_phantom(81, 0.25, rods=False)
    g = scenes.mc_geom(33, 32.5 / 33, n_views=3)
    g.angle_step_deg = 37.0
    vol = scenes.volume_for(lab, 0.25)
    xs = scenes.make_xs()
    m = oracle.project_primary(g, vol, lab, oracle.tables_from_xs(xs), 140.0)
    mu = float(xs.total[0][140])
    for v in range(3):
        i = j = 16                                           # central pixel: chord = diameter 20 cm
        assert abs(m[v, i, j] / mu - 20.0) < 0.5
    assert m[:, 0, :].max() == 0                            # the edge rays miss the cylinder
    assert (m >= 0).all()


def test_counts_to_map_known_answers(oracle):
    c = np.array([0, 1, 50, 2000, 5000], np.int32)
    m = oracle.counts_to_map(c, 2000)                       # CBCT_real325im.cu:267-285
    assert m[0] == m[1] and abs(m[0] - math.log(2000.0)) < 1e-6       # 0 counts are clamped to 1
    assert abs(m[2] - math.log(40.0)) < 1e-6
    # -log(int) is the double overload, log(float(per)) the float one: counts == per gives the
    # float rounding of log(2000), not exactly 0; counts > per are clamped to per
    assert m[3] == m[4] and abs(m[3]) < 2e-7


def test_energy_integrating_detector_mode(oracle):
    """SURVEY 8f-3, off by default: with detector_mode = ENERGY every detected photon adds
    (int)(16 E + 0.5) instead of 1 (CBCT_real325im.cu:589-590 counts photons).  Same histories, so the
    primary image is the count image times 16*140 and the scatter part sums to the detected scatter energy."""
    lab = scenes.cylinder_phantom(33, 1.0)
    g = scenes.mc_geom(9, 32.5 / 9, n_views=1)
    vol = scenes.volume_for(lab, 1.0)
    tb = oracle.tables_from_xs(scenes.make_xs())
    opts = oracle.mc_opts(oracle.RNG_PHILOX, seed=4)
    c0, c5, rc, _, _ = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), opts, 300)
    g.detector_mode = 1
    e0, e5, re, _, _ = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(140.0), opts, 300)
    assert rc["primaries"] == re["primaries"] and rc["scatter_detected"] == re["scatter_detected"] > 0
    assert np.array_equal(e0, c0 * (16 * 140))
    scat = (e5.astype(np.int64) - e0).sum()
    assert abs(scat / 16.0 - re["sum_e_scatter"]) <= 0.5 / 16 * re["scatter_detected"] + 1e-6
    assert scat < 16 * 140 * re["scatter_detected"]            # Compton-scattered photons arrive with less than 140 keV


@pytest.mark.parametrize("keV,x0", [(30.0, 1.0), (80.0, 2.2), (140.0, 0.9)])
def test_rayleigh_formfactor_sampler_against_closed_form(oracle, keV, x0):
    """SURVEY 8f-3 (coherent_mode FORMFACTOR; not in the reference): the accepted cos(theta) of the table sampler
    follow p(c) ~ (1 + c^2) F(x)^2, x^2 = x2max (1 - c)/2, for the analytic F^2 = (1 + x^2/x0^2)^-4 the tables were
    built from (monte_xs_formfactor_hydrogenic)."""
    from monte_b200 import api
    xs = scenes.make_xs()
    assert api.load().monte_xs_formfactor_hydrogenic(xs, 0, x0) == 0
    api.load().monte_xs_formfactor_hydrogenic(xs, 1, 1.0)
    assert xs.ff_points == 128 and xs.ff_x2[0][0] == 0.0 and xs.ff_x2[0][127] >= (200 / 12.3984) ** 2
    tb = oracle.tables_from_xs(xs)
    rng = np.random.default_rng(int(keV))
    n = 60000
    u = rng.random((n, 2))
    cs = np.array([c for ok, c in (oracle.rayleigh_round(tb, 0, keV, a, b) for a, b in u) if ok])
    assert 0.5 * n <= cs.size <= n                                     # acceptance (1 + c^2)/2 >= 1/2
    x2max = (keV / 12.3984193) ** 2
    edges = np.linspace(-1.0, 1.0, 21)
    fine = np.linspace(-1.0, 1.0, 200001)
    pdf = (1 + fine ** 2) * (1 + x2max * (1 - fine) / 2 / x0 ** 2) ** -4.0
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(fine))])
    expect = np.diff(np.interp(edges, fine, cdf)) / cdf[-1] * cs.size
    got = np.histogram(cs, edges)[0]
    big = expect > 20
    chi2 = ((got[big] - expect[big]) ** 2 / expect[big]).sum()
    dof = int(big.sum()) - 1
    assert chi2 < dof + 6 * math.sqrt(2 * dof) + 10, (chi2, dof, got, expect)   # + table-interpolation bias (128 points)
    # forward peaked, and the more so the higher the energy / the more diffuse the charge cloud
    assert cs.mean() > 0.2


def test_rayleigh_mode_changes_only_coherent_histories(oracle):
    """with the form factor on, every history without a coherent event keeps its fate; energy is unchanged by
    coherent events (elastic): detected single-coherent photons still carry the source energy"""
    lab = scenes.cylinder_phantom(33, 1.0)
    g = scenes.mc_geom(9, 32.5 / 9, n_views=1)
    vol = scenes.volume_for(lab, 1.0)
    xs = scenes.add_formfactors(scenes.make_xs())
    tb = oracle.tables_from_xs(xs)
    opts = oracle.mc_opts(oracle.RNG_PHILOX, seed=4)
    a0, a5, ra, fa, ea = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(40.0), opts, 400, want_fates=True)
    g.coherent_mode = 1
    b0, b5, rb, fb, eb = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(40.0), opts, 400, want_fates=True)
    assert np.array_equal(a0, b0) and ra["primaries"] == rb["primaries"]
    changed = fa != fb
    assert 0.01 < changed.mean() < 0.3
    assert rb["coherent"] > 0 and rb["scatter_detected"] < ra["scatter_detected"]    # deflected photons mostly miss
    # MT19937 mode runs too and agrees statistically on the totals
    g.coherent_mode = 1
    c0, c5, rc, _, _ = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(40.0), oracle.mc_opts(oracle.RNG_MT, seed=3), 400)
    for k in ("absorbed", "coherent", "compton"):
        assert abs(rc[k] - rb[k]) < 6 * math.sqrt(rc[k] + rb[k] + 1), (k, rc[k], rb[k])


def test_clearance_grid_is_a_safe_lower_bound():
    """monte_mc_clearance_grid (host helper): for every cell, no voxel of the heavy material lies closer to any
    point of the cell than grid * half a cell side -- brute force on a small volume -- and the bound is tight to
    within one grid unit somewhere"""
    from monte_b200 import api
    rng = np.random.default_rng(5)
    lab = np.zeros((20, 24, 28), np.uint8)
    lab[2:18, 3:20, 4:25] = 1
    for _ in range(6):                                   # a few heavy blobs
        z, y, x = rng.integers(2, 18), rng.integers(3, 20), rng.integers(4, 25)
        lab[z:z + 2, y:y + 1, x:x + 3] = 2
    lab[10, 11, 12] = 7                                  # label above n_materials clamps to the last (= heavy) material
    vol = scenes.volume_for(lab, 0.5, tight=False)
    xs = scenes.make_xs()
    for cl in (0, 1, 2):
        grid, heavy = api.clearance_grid(vol, lab, xs, cell_log2=cl)
        assert heavy == 1
        c = 1 << cl
        hz, hy, hx = np.nonzero(lab >= 2)
        gz, gy, gx = grid.shape
        assert (gz, gy, gx) == (-(-20 // c), -(-24 // c), -(-28 // c))
        slack = []
        for cz in range(gz):
            for cy in range(gy):
                for cx in range(gx):
                    # smallest distance (voxel units) between the cell's box and any heavy VOXEL's box
                    dz = np.maximum(0, np.maximum(cz * c - (hz + 1), hz - (cz + 1) * c))
                    dy = np.maximum(0, np.maximum(cy * c - (hy + 1), hy - (cy + 1) * c))
                    dx = np.maximum(0, np.maximum(cx * c - (hx + 1), hx - (cx + 1) * c))
                    true = np.sqrt(dz * dz + dy * dy + dx * dx).min()
                    reach = grid[cz, cy, cx] * 0.5 * c
                    assert reach <= true + 1e-9, (cl, cz, cy, cx, reach, true)
                    slack.append(true - reach)
        assert min(slack) < 0.5 * c + 1e-9 and np.mean(slack) < 1.5 * c      # conservative by at most about a cell


def test_clearance_tracking_is_the_same_physics(oracle):
    """tracking_mode CLEARANCE against the reference's single-majorant loop, both on MT19937: images agree by
    chi-square, totals within 5 sigma, with a fraction of the tentative collisions"""
    from monte_b200 import api
    lab = scenes.cylinder_phantom(41, 0.5)
    g = scenes.mc_geom(13, 32.5 / 13, n_views=2)
    g.angle_step_deg = 22.5
    vol = scenes.volume_for(lab, 0.5)
    xs = scenes.make_xs()
    tb = oracle.tables_from_xs(xs)
    spec, keep = scenes.kramers_spectrum()
    per = 3000
    a0, a5, ra, _, _ = oracle.mc_run(g, vol, lab, tb, spec, oracle.mc_opts(oracle.RNG_MT, seed=1), per)
    vol.tracking_mode, vol.clearance_cell_log2 = 1, 1
    grid, heavy = api.clearance_grid(vol, lab, xs)
    o, keepg = oracle.with_clearance(oracle.mc_opts(oracle.RNG_MT, seed=2), grid, heavy)
    b0, b5, rb, _, _ = oracle.mc_run(g, vol, lab, tb, spec, o, per)
    assert rb["woodcock_steps"] < 0.45 * ra["woodcock_steps"]
    for k in ("primaries", "scatter_detected", "absorbed", "compton", "coherent", "interactions"):
        assert abs(ra[k] - rb[k]) < 5 * math.sqrt(ra[k] + rb[k] + 1), (k, ra[k], rb[k])
    for x, y, binom in ((a0, b0, per), (a5 - a0, b5 - b0, None)):
        d = x.astype(np.float64) - y
        var = (x + y).astype(np.float64)
        if binom:
            var = var * (1.0 - (x + y) / (2.0 * binom))
        m = var > 0.5
        c2, dof = (d[m] ** 2 / var[m]).sum(), int(m.sum())
        assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), (c2, dof)
    assert abs(ra["sum_e_scatter"] / ra["scatter_detected"] - rb["sum_e_scatter"] / rb["scatter_detected"]) < 3.0


def test_auto_tracking_resolution_follows_the_measured_crossover():
    """MONTE_MC_TRACK_AUTO (host helper, deterministic): single majorant at 140 keV, the directional two-level majorant at
    60 keV and for the 120 kVp spectrum -- the side of the crossover each was measured on"""
    from monte_b200 import api
    xs = scenes.make_xs()
    mode, cl, r = api.resolve_tracking(xs, scenes.mono_spectrum(140.0))
    assert mode == _abi.TRACK_GLOBAL and 1.5 < r < 2.2
    mode, cl, r = api.resolve_tracking(xs, scenes.mono_spectrum(60.0))
    assert mode == _abi.TRACK_DIRECTIONAL and cl == 2 and 4.0 < r < 6.0
    spec, keep = scenes.kramers_spectrum()
    mode, cl, r = api.resolve_tracking(xs, spec)
    assert mode == _abi.TRACK_DIRECTIONAL and r > 4.0
    assert api.resolve_tracking(scenes.make_xs(("h2o",)), spec)[0] == _abi.TRACK_GLOBAL      # nothing to exclude
