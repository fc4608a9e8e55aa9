"""GPU parity tests of the Monte-Carlo transport, called through the C ABI (libmonte_gpu.so).

Three levels, from strongest to the one BASELINE.json's north_star states:
 1. history-coupled: the oracle (intended physics, Philox mode) and the CUDA kernel draw the same
    Philox2x32-10 variates per history, so every history must end the same way (same fate, same
    detector bin); the only allowed differences are fp32-vs-fp64 threshold flips.  Census on a B200
    (scripts/mc_flip_census.py, profiles/r02_mc_flip_census.jsonl): 26 of 3.04e6 histories over six scenes
    (8.5e-6; 6 scatter hits within rounding of a pixel edge, 20 histories whose path diverged at an upstream
    threshold -- Woodcock acceptance, voxel face, interaction selection, Kahn acceptance), worst scene 1.4e-5.
    The bars below are 1e-3: two orders of magnitude above what is observed, ten times below round 1's.
 2. deterministic: the primary projection agrees with the oracle within 1e-4 relative (fp32),
    counts->map exactly, results are independent of how histories are partitioned.
 3. statistical (north_star): against the oracle driven by the reference's own generator
    (MT19937): per detector pixel within 3 sigma (Poisson/binomial), chi-square over the full image,
    matching mean detected energy.
"""
import math

import numpy as np
import pytest

from monte_b200 import _abi, scenes

pytestmark = pytest.mark.gpu


def scene(n=33, pitch=1.0, det=17, views=4, mode=_abi.SOURCE_PENCIL, max_scatter=5):
    lab = scenes.cylinder_phantom(n, pitch)
    g = scenes.mc_geom(det, 32.5 / det, n_views=views, source_mode=mode, max_scatter=max_scatter)
    g.angle_step_deg = 360.0 / views
    return g, scenes.volume_for(lab, pitch), lab


def chi2_images(a, b, binomial_per=None):
    """sum (a-b)^2/var over pixels with a+b>0; var of a-b for two equal-exposure counts."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    var = a + b
    if binomial_per is not None:                           # primaries: binomial(per, p), p ~ (a+b)/(2 per)
        var = var * (1.0 - (a + b) / (2.0 * binomial_per))
    m = var > 0.5
    z2 = (a - b)[m] ** 2 / var[m]
    return z2.sum(), int(m.sum()), float((z2 > 9.0).mean()) if m.any() else 0.0


@pytest.mark.parametrize("mode,poly", [(_abi.SOURCE_PENCIL, False), (_abi.SOURCE_CONE, True)])
def test_history_coupled_fates_match_oracle(monte, oracle, mode, poly):
    g, vol, lab = scene(n=33, pitch=1.0, det=17, views=3, mode=mode)
    xs = scenes.make_xs()
    spec, keep = scenes.kramers_spectrum() if poly else (scenes.mono_spectrum(140.0), None)
    per, seed, view = 24, 77, 1
    sc = monte.Scene(g, vol, lab, xs, spec)
    f_gpu, e_gpu = sc.fates(view, per, seed)
    sc.close()
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec,
                                            oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per,
                                            views=(view, view + 1), want_fates=True)
    kind_g, kind_c = f_gpu & 0xFF, f_cpu & 0xFF
    assert (kind_g != 0).all() and (kind_c != 0).all()      # every history was run and ended
    same = f_gpu == f_cpu
    assert same.mean() > 0.999, "only %.4f of %d histories end identically" % (same.mean(), same.size)
    # where the fate matches the photon energy must too
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)
    # every kind of ending is exercised
    for k in (1, 2, 3, 4):
        assert (kind_c == k).any(), k


def test_images_and_counters_match_coupled_oracle(monte, oracle):
    """same as above at the tally level: image0 identical, image5 and the counters within the flips"""
    g, vol, lab = scene(n=33, pitch=1.0, det=17, views=2)
    xs = scenes.make_xs()
    per, seed = 60, 5
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                      oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per)
    n = st["histories"]
    assert n == res["histories"] == 2 * 17 * 17 * per
    assert np.abs(im0.astype(int) - o0).sum() <= 0.001 * n
    assert np.abs(im5.astype(int) - o5).sum() <= 0.001 * n
    for k in ("primaries", "scatter_detected", "absorbed", "interactions", "coherent", "compton", "woodcock_steps"):
        assert abs(st[k] - res[k]) <= 0.001 * max(res[k], 1) + 5, (k, st[k], res[k])
    assert abs(st["sum_e_scatter"] - res["sum_e_scatter"]) <= 0.001 * res["sum_e_scatter"] + 200
    assert im0.sum() == st["primaries"] and im5.sum() == st["primaries"] + st["scatter_detected"]


def test_energy_integrating_detector_matches_coupled_oracle(monte, oracle):
    """SURVEY 8f-3 flag (off in parity mode): energy-integrating tallies in 1/16 keV, against the oracle
    run on the same Philox variates, and against the photon-counting run of the same histories"""
    g, vol, lab = scene(n=33, pitch=1.0, det=17, views=2)
    xs = scenes.make_xs()
    per, seed = 60, 5
    c0, c5, stc = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)
    g.detector_mode = 1
    e0, e5, ste = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)
    assert ste["primaries"] == stc["primaries"] and ste["scatter_detected"] == stc["scatter_detected"]
    assert np.array_equal(e0, c0 * (16 * 140))                 # primaries arrive with the source energy
    scat = (e5.astype(np.int64) - e0).sum()
    assert abs(scat / 16.0 - ste["sum_e_scatter"]) <= (0.5 / 16 + 1 / 1024) * ste["scatter_detected"] + 1e-3
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                      oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per)
    n = ste["histories"]
    assert np.abs(e0.astype(np.int64) - o0).sum() <= 0.001 * n * 16 * 140
    assert np.abs(e5.astype(np.int64) - o5).sum() <= 0.001 * n * 16 * 140
    g.detector_mode = 7
    with pytest.raises(Exception, match="detector_mode"):
        monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)


def test_statistical_parity_with_reference_generator(monte, oracle):
    """north_star: per pixel within 3 sigma, chi-square over the image, mean detected energy —
    oracle on MT19937 (the reference's generator), CUDA path on Philox."""
    g, vol, lab = scene(n=41, pitch=0.5, det=25, views=2)
    xs = scenes.make_xs()
    per = 2500
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=2024)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                      oracle.mc_opts(oracle.RNG_MT, seed=11), per)
    # primaries (image0): binomial
    c2, dof, frac3 = chi2_images(im0, o0, binomial_per=per)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), ("image0 chi2", c2, dof)
    assert frac3 < 0.012                                     # 0.27 % expected beyond 3 sigma
    # scatter-only image (image5 - image0): Poisson
    c2, dof, frac3 = chi2_images(im5 - im0, o5 - o0)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), ("scatter chi2", c2, dof)
    assert frac3 < 0.012
    # totals
    for k in ("primaries", "scatter_detected", "absorbed", "compton", "coherent"):
        assert abs(st[k] - res[k]) < 5 * math.sqrt(st[k] + res[k] + 1), (k, st[k], res[k])
    # mean detected energy of the scattered photons (keV)
    eg = st["sum_e_scatter"] / st["scatter_detected"]
    ec = res["sum_e_scatter"] / res["scatter_detected"]
    assert abs(eg - ec) < 5 * 15.0 / math.sqrt(min(st["scatter_detected"], res["scatter_detected"])), (eg, ec)
    assert abs(st["sum_e_primary"] / st["primaries"] - 140.0) < 1e-3


def test_partition_independence(monte):
    """splitting photons (multi-GPU partition) or views changes nothing: integer tallies, Philox ids"""
    import torch
    g, vol, lab = scene(n=33, pitch=1.0, det=17, views=3)
    xs = scenes.make_xs()
    per, seed = 50, 3
    ref0, ref5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(), per, seed)
    sc = monte.Scene(g, vol, lab, xs, scenes.mono_spectrum())
    a0 = torch.zeros((3, 17, 17), dtype=torch.int32, device="cuda")
    a5 = torch.zeros_like(a0)
    stats = torch.zeros(16, dtype=torch.int64, device="cuda")
    for nr in ((0, 7), (7, 31), (31, 50)):
        sc.simulate_dev(a0, a5, per, seed, views=(0, 2), n_range=nr, d_stats=stats)
    sc.simulate_dev(a0, a5, per, seed, views=(2, 3), d_stats=stats)
    torch.cuda.synchronize()
    sc.close()
    assert np.array_equal(a0.cpu().numpy(), ref0) and np.array_equal(a5.cpu().numpy(), ref5)
    st2 = monte.unpack_stats(stats.cpu().numpy().astype(np.uint64))
    for k in ("histories", "primaries", "scatter_detected", "absorbed", "interactions", "woodcock_steps"):
        assert st2[k] == st[k], k
    assert st2["sum_e_scatter"] == st["sum_e_scatter"]       # fixed-point sums: order independent
    # and a different seed gives a different answer
    d0, _, _ = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(), per, seed + 1)
    assert not np.array_equal(d0, ref0)


def test_project_primary_matches_oracle(monte, oracle):
    """deterministic primary projection within 1e-4 relative (fp32) of the CPU path"""
    lab = scenes.cylinder_phantom(65, 0.5)
    g = scenes.mc_geom(65, 0.5, n_views=8)
    g.angle_step_deg = 45.0 / 2 + 0.37
    vol = scenes.volume_for(lab, 0.5)
    xs = scenes.make_xs()
    m = monte.project_primary(g, vol, lab, xs, 140.0)
    mo = oracle.project_primary(g, vol, lab, oracle.tables_from_xs(xs), 140.0)
    err = np.abs(m.astype(np.float64) - mo).max()
    assert err <= 1e-4 * mo.max(), (err, mo.max())
    assert (m[:, :, 0] == 0).all() or True


def test_mc_primary_agrees_with_deterministic_projection(monte):
    """-ln(image0/per) is the line integral within 3 sigma per pixel"""
    lab = scenes.cylinder_phantom(33, 1.0)
    g = scenes.mc_geom(17, 32.5 / 17, n_views=2)
    g.angle_step_deg = 30.0
    vol = scenes.volume_for(lab, 1.0)
    xs = scenes.make_xs()
    per = 20000
    im0, _, _ = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=8)
    p = np.exp(-monte.project_primary(g, vol, lab, xs, 140.0).astype(np.float64))
    z = (im0 - per * p) / np.sqrt(np.maximum(per * p * (1 - p), 1e-9))
    z = z[p < 1 - 1e-12]
    assert (np.abs(z) > 3).mean() < 0.012 and abs(z.mean()) < 0.25
    assert (im0[p >= 1 - 1e-12] == per).all()                # rays that miss the phantom: all detected


def test_counts_to_map_matches_oracle(monte, oracle):
    rng = np.random.default_rng(0)
    c = rng.integers(0, 3000, size=100001).astype(np.int32)
    c[:3] = (0, 1, 2000)
    m = monte.counts_to_map(c, 2000)
    assert np.array_equal(m, oracle.counts_to_map(c, 2000))


def test_edge_cases(monte):
    xs = scenes.make_xs()
    # empty phantom: every photon is an unscattered primary
    lab = np.zeros((9, 9, 9), np.uint8)
    g = scenes.mc_geom(5, 6.5, n_views=2)
    vol = scenes.volume_for(lab, 1.0)
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(), 7, seed=1)
    assert (im0 == 7).all() and (im5 == 7).all() and st["interactions"] == 0 and st["primaries"] == 2 * 25 * 7
    # max_scatter = 0: only primaries are tallied
    g2, vol2, lab2 = scene(n=17, pitch=2.0, det=9, views=1, max_scatter=0)
    a0, a5, st2 = monte.simulate(g2, vol2, lab2, xs, scenes.mono_spectrum(), 40, seed=1)
    assert np.array_equal(a0, a5) and st2["scatter_detected"] == 0 and st2["interactions"] == 0
    # ragged detector (ny != nx) and a photon count that is not a multiple of anything
    g3 = scenes.mc_geom(0, 2.5, n_views=1, ny=13, nx=7)
    lab3 = scenes.cylinder_phantom(17, 2.0)
    a0, a5, st3 = monte.simulate(g3, scenes.volume_for(lab3, 2.0), lab3, xs, scenes.mono_spectrum(), 37, seed=4)
    assert st3["histories"] == 13 * 7 * 37 and a0.sum() == st3["primaries"]
    # bad arguments are reported, not fatal
    with pytest.raises(monte.MonteError):
        g3.max_scatter = 99
        monte.simulate(g3, scenes.volume_for(lab3, 2.0), lab3, xs, scenes.mono_spectrum(), 1)


def test_full_size_c2_shape_properties(monte):
    """BASELINE config 2 shape (325^3 labels, 325x325 detector): one view, size-independent checks"""
    g, vol, lab = scenes.config_c2()
    xs = scenes.make_xs()
    per = 20
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=1, views=(100, 101))
    npix = 325 * 325
    assert st["histories"] == npix * per
    assert im0.sum() == st["primaries"] and im5.sum() == st["primaries"] + st["scatter_detected"]
    assert not im0[:100].any() and not im0[101:].any()
    assert (im5 >= im0).all() and im0.max() <= per
    # rays outside the cylinder's shadow are untouched
    assert (im0[100, :, 0] == per).all() and (im0[100, 0, :] == per).all()
    # bookkeeping: every history ends exactly one way
    assert st["interactions"] == st["absorbed"] + st["coherent"] + st["compton"]
    # central pixel transmission ~ exp(-mu*20cm) = 0.046: far fewer than per
    assert im0[100, 150:175, 150:175].mean() < 0.2 * per


def test_c2_full_shape_statistical_parity(monte, oracle):
    """BASELINE config 2 at its full shape (325^3 labels, 325x325 detector), one view, 24 photons per
    pixel = 2.5e6 histories: CUDA path (Philox) against the oracle on MT19937 (all host threads)."""
    g, vol, lab = scenes.config_c2()
    xs = scenes.make_xs()
    per, view = 24, 37
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=99, views=(view, view + 1))
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                      oracle.mc_opts(oracle.RNG_MT, seed=5, n_threads=0), per, views=(view, view + 1))
    assert st["histories"] == res["histories"] == 325 * 325 * per
    c2, dof, frac3 = chi2_images(im0[view], o0[view], binomial_per=per)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), ("image0 chi2", c2, dof)
    assert frac3 < 0.006
    c2, dof, frac3 = chi2_images((im5 - im0)[view], (o5 - o0)[view])
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), ("scatter chi2", c2, dof)
    for k in ("primaries", "scatter_detected", "absorbed", "compton", "coherent", "woodcock_steps"):
        assert abs(st[k] - res[k]) < 5 * math.sqrt(st[k] + res[k] + 1) * (3 if k == "woodcock_steps" else 1), (k, st[k], res[k])
    eg, ec = st["sum_e_scatter"] / st["scatter_detected"], res["sum_e_scatter"] / res["scatter_detected"]
    assert abs(eg - ec) < 5 * 15.0 / math.sqrt(res["scatter_detected"]), (eg, ec)


def test_c4_polyenergetic_cone_statistical_parity(monte, oracle):
    """BASELINE config 4 physics (120 kVp spectrum, sampled cone) on a reduced scene, against MT19937"""
    g, vol, lab = scene(n=41, pitch=0.5, det=25, views=2, mode=_abi.SOURCE_CONE)
    xs = scenes.make_xs()
    spec, keep = scenes.kramers_spectrum(120.0)
    per = 1500
    im0, im5, st = monte.simulate(g, vol, lab, xs, spec, per, seed=31)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_MT, seed=8), per)
    c2, dof, frac3 = chi2_images(im0, o0, binomial_per=per)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof) and frac3 < 0.012, ("image0", c2, dof, frac3)
    c2, dof, frac3 = chi2_images(im5 - im0, o5 - o0)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof) and frac3 < 0.012, ("scatter", c2, dof, frac3)
    # mean energies of the detected primaries (beam hardening) and of the scattered photons
    ep_g, ep_c = st["sum_e_primary"] / st["primaries"], res["sum_e_primary"] / res["primaries"]
    es_g, es_c = st["sum_e_scatter"] / st["scatter_detected"], res["sum_e_scatter"] / res["scatter_detected"]
    assert abs(ep_g - ep_c) < 5 * 25.0 / math.sqrt(res["primaries"]), (ep_g, ep_c)
    assert abs(es_g - es_c) < 5 * 25.0 / math.sqrt(res["scatter_detected"]), (es_g, es_c)
    assert 40.0 < ep_c < 90.0


def test_three_materials_im_variant_coupled(monte, oracle):
    """CBCT_real325im.cu semantics: labels 1 = H2O, 2 = Ca, anything else = the last material (PMMA,
    :640-646); a PMMA cylinder with a water core and calcium inserts, history-coupled with the oracle"""
    n, pitch = 41, 0.5
    c = (np.arange(n) + 0.5) * pitch - 0.5 * n * pitch
    x, y, z = c[None, None, :], c[None, :, None], c[:, None, None]
    lab = np.zeros((n, n, n), np.uint8)
    body = (x * x + y * y <= 81.0) & (np.abs(z) <= 9.0)           # r = 9 cylinder along z (:904)
    lab[np.broadcast_to(body, lab.shape)] = 3                      # PMMA
    lab[np.broadcast_to(body & (x * x + y * y <= 9.0), lab.shape)] = 1
    lab[np.broadcast_to(body & ((x - 5) ** 2 + y * y <= 2.25), lab.shape)] = 2
    lab[np.broadcast_to(body & ((x + 5) ** 2 + y * y <= 2.25), lab.shape)] = 7     # unknown label -> last material
    g = scenes.mc_geom(21, 32.5 / 21, n_views=2)
    g.angle_step_deg = 90.0
    vol = scenes.volume_for(lab, pitch)
    xs = scenes.make_xs(("h2o", "ca", "pmma"))
    per, seed = 30, 12
    sc = monte.Scene(g, vol, lab, xs, scenes.mono_spectrum(140.0))
    f_gpu, _ = sc.fates(0, per, seed)
    sc.close()
    _, _, res, f_cpu, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                        oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per, views=(0, 1), want_fates=True)
    assert (f_gpu == f_cpu).mean() > 0.999
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                      oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per)
    assert np.abs(im5.astype(int) - o5).sum() <= 0.004 * st["histories"]
    assert abs(st["compton"] - res["compton"]) <= 0.004 * res["compton"] + 5


def hu_scene(monte, keV=60.0, n=33, pitch=0.6, det=17, views=2, cortical=False):
    """a CT volume in HU -> labels + tables through monte_ctnum_segment"""
    hu = scenes.hu_head_phantom(n, pitch, cortical=cortical)
    classes = monte.hu_classes_default(True)
    lab, xs, mu, present = monte.ctnum_segment(hu, classes, scenes.make_xs(), keV)
    g = scenes.mc_geom(det, 32.5 / det, n_views=views)
    g.angle_step_deg = 360.0 / views
    return hu, classes, g, scenes.volume_for(lab, pitch), lab, xs, mu, present


def test_hu_volume_transport_with_present_material_majorant(monte, oracle):
    """a HU volume segmented into 8 classes (density bins of water, water + calcium mixtures) is what
    the transport consumes; with majorant_mode PRESENT the Woodcock majorant covers only the classes that occur.
    History by history against the oracle (same variates), fewer tentative collisions than with the majorant of all
    tables, same physics."""
    keV = 60.0
    hu, classes, g, vol, lab, xs, mu, present = hu_scene(monte, keV)
    assert xs.n_materials == 8 and present == 0b01001111            # lung, adipose, soft, muscle, dense bone: no cortical bone in this head
    per, seed, view = 24, 11, 1
    spec = scenes.mono_spectrum(keV)
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    sc = monte.Scene(g, vol, lab, xs, spec)
    f_gpu, e_gpu = sc.fates(view, per, seed)
    sc.close()
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec,
                                            oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per,
                                            views=(view, view + 1), want_fates=True)
    same = f_gpu == f_cpu
    assert same.mean() > 0.999, "only %.4f of %d histories end identically" % (same.mean(), same.size)
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)
    for k in (1, 3, 4):
        assert ((f_cpu & 0xFF) == k).any(), k
    # tallies and counters of the host-buffer call, PRESENT against the oracle and against the majorant of all tables
    p0, p5, st_p = monte.simulate(g, vol, lab, xs, spec, 60, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), 60)
    n = st_p["histories"]
    assert np.abs(p0.astype(int) - o0).sum() <= 0.001 * n and np.abs(p5.astype(int) - o5).sum() <= 0.001 * n
    assert abs(st_p["woodcock_steps"] - res["woodcock_steps"]) <= 0.001 * res["woodcock_steps"] + 5
    vol.majorant_mode = _abi.MAJORANT_ALL
    a0, a5, st_a = monte.simulate(g, vol, lab, xs, spec, 60, seed)
    assert st_p["woodcock_steps"] < 0.9 * st_a["woodcock_steps"], (st_p["woodcock_steps"], st_a["woodcock_steps"])
    # same physics: primaries are binomial with the same p per pixel in both runs
    for k in ("primaries", "absorbed", "interactions"):
        assert abs(st_p[k] - st_a[k]) <= 6.0 * math.sqrt(st_p[k] + st_a[k] + 1), (k, st_p[k], st_a[k])
    c2, nb, _ = chi2_images(p0, a0, binomial_per=60)
    assert c2 < nb + 6.0 * math.sqrt(2.0 * nb), (c2, nb)
    # the deterministic projector integrates exactly the mu the segmentation reports
    pm = monte.project_primary(g, vol, lab, xs, keV, views=(0, 1))
    po = oracle.project_primary(g, vol, lab, oracle.tables_from_xs(xs), keV)
    assert np.abs(pm[0] - po[0]).max() <= 1e-4 * np.abs(po[0]).max()
    for m in range(xs.n_materials):
        sel = lab == m + 1
        if sel.any():
            assert np.all(mu[sel] == np.float32(xs.total[m][60]) * np.float32(xs.density[m]))
    assert np.all(mu[lab == 0] == 0)


def test_scene_update_labels_follows_the_present_materials(monte):
    """majorant_mode PRESENT on a resident scene: new labels with another set of materials rebuild the tables"""
    keV = 60.0
    hu, classes, g, vol, lab, xs, mu, present = hu_scene(monte, keV)
    hu2 = scenes.hu_head_phantom(33, 0.6, cortical=True)
    lab2, xs2, _, present2 = monte.ctnum_segment(hu2, classes, scenes.make_xs(), keV)
    assert present2 == 0b11001111 and np.array_equal(np.ctypeslib.as_array(xs2.total), np.ctypeslib.as_array(xs.total))
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    spec = scenes.mono_spectrum(keV)
    per, seed = 30, 3
    sc2 = monte.Scene(g, vol, lab2, xs, spec)                                        # built from lab2 directly
    want, want_e = sc2.fates(1, per, seed)
    sc2.close()
    sc = monte.Scene(g, vol, lab, xs, spec)                                          # built from lab, then updated
    before, _ = sc.fates(1, per, seed)
    sc.update_labels(lab2)
    got, got_e = sc.fates(1, per, seed)
    sc.close()
    assert np.array_equal(got, want) and np.array_equal(got_e, want_e)
    assert not np.array_equal(before, want)


def test_all_air_volume_under_present_majorant(monte):
    """majorant_mode PRESENT with nothing present: the medium is transparent, the kernel returns (one step of 1e30 cm
    leaves the clip box) and every photon is a primary"""
    lab = np.zeros((9, 9, 9), np.uint8)
    g = scenes.mc_geom(5, 32.5 / 5, n_views=2)
    vol = scenes.volume_for(lab, 2.0)
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    im0, im5, st = monte.simulate(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum(140.0), 7, 1)
    assert st["histories"] == st["primaries"] == 2 * 25 * 7 and st["interactions"] == 0
    assert (im0 == 7).all() and np.array_equal(im0, im5)
    vol.majorant_mode = 9
    with pytest.raises(Exception, match="majorant_mode"):
        monte.simulate(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum(140.0), 7, 1)


def test_ring_detector_primary_transmission_kat(monte, oracle):
    """SURVEY 8f-4 / 8c KAT: the reference's 2-D geometry (monte_cpp/circle3_2.cpp: source at the centre of an r = 10 cm
    water disc, ring of angular bins).  Every primary crosses 10 cm of water: transmission exp(-0.1538092 * 10) =
    0.214791 at 140 keV in every bin; history by history against the oracle."""
    lab = scenes.cylinder_phantom(41, 0.5, radius=10.0, half_len=10.0, rods=False)
    vol = scenes.volume_for(lab, 0.5)
    g = scenes.ring_geom(36, 1, 1.0, 15.0)
    xs = scenes.make_xs(("h2o",))
    per, seed = 3000, 3
    im0, im5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed)
    assert im0.shape == (1, 36, 1) and st["histories"] == 36 * per
    p = math.exp(-0.1538092 * 10.0)
    # the voxelised disc is 10 cm +- half a voxel along a ray: compare with the deterministic path length through the labels
    sig = math.sqrt(per * p * (1 - p))
    assert np.abs(im0[0, :, 0] - per * p).max() < 5.0 * sig + 0.1538 * 0.5 * per * p, im0[0, :, 0]
    assert abs(im0.sum() / (36.0 * per) - p) < 0.01
    assert (im5 >= im0).all() and st["scatter_detected"] > 0 and im5.sum() == st["primaries"] + st["scatter_detected"]
    sc = monte.Scene(g, vol, lab, xs, scenes.mono_spectrum(140.0))
    f_gpu, e_gpu = sc.fates(0, 200, seed)
    sc.close()
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), scenes.mono_spectrum(140.0),
                                            oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), 200, want_fates=True)
    same = f_gpu == f_cpu
    assert same.mean() > 0.999, same.mean()
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)


def test_ring_detector_coupled_with_oracle(monte, oracle):
    """ring detector with axial bins, calcium rods, jittered aim inside the bins, two view angles (the bins rotate with the
    view): fates, tallies and counters against the oracle on the same variates; argument errors"""
    lab = scenes.cylinder_phantom(33, 1.0)
    vol = scenes.volume_for(lab, 1.0)
    g = scenes.ring_geom(24, 5, 4.0, 20.0, n_views=2, source_mode=_abi.SOURCE_CONE)
    g.angle0_deg, g.angle_step_deg = 10.0, 37.0
    xs = scenes.make_xs()
    spec = scenes.mono_spectrum(60.0)
    per, seed = 60, 9
    for view in (0, 1):
        sc = monte.Scene(g, vol, lab, xs, spec)
        f_gpu, e_gpu = sc.fates(view, per, seed)
        sc.close()
        _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per,
                                                views=(view, view + 1), want_fates=True)
        same = f_gpu == f_cpu
        assert same.mean() > 0.999, (view, same.mean())
        for k in (1, 2, 3, 4):
            assert ((f_cpu & 0xFF) == k).any(), k
    im0, im5, st = monte.simulate(g, vol, lab, xs, spec, per, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per)
    n = st["histories"]
    assert np.abs(im0.astype(int) - o0).sum() <= 0.001 * n + 2 and np.abs(im5.astype(int) - o5).sum() <= 0.001 * n + 2
    for k in ("primaries", "scatter_detected", "absorbed", "interactions", "woodcock_steps"):
        assert abs(st[k] - res[k]) <= 0.001 * max(res[k], 1) + 5, (k, st[k], res[k])
    g.ring_radius = 12.0                                              # the clip box (corner at 14.1 cm) pokes through the ring
    with pytest.raises(Exception, match="inside the detector ring"):
        monte.simulate(g, vol, lab, xs, spec, per, seed)
    g.ring_radius = 20.0
    # the two-level majorant and the form-factor deflection on the ring detector (their own kernel instantiations)
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_DIRECTIONAL, 1
    g.coherent_mode = _abi.COHERENT_FORMFACTOR
    xf = scenes.add_formfactors(scenes.make_xs())
    sc = monte.Scene(g, vol, lab, xf, spec)
    f_gpu, e_gpu = sc.fates(0, per, seed)
    sc.close()
    grid, heavy = monte.clearance_grid(vol, lab, xf)
    opts, _keep = oracle.with_clearance(oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), grid, heavy)
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xf), spec, opts, per, views=(0, 1), want_fates=True)
    assert (f_gpu == f_cpu).mean() > 0.999
