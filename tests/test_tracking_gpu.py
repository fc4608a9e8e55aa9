"""GPU parity tests of the two-level Woodcock majorant (monte_mc_volume.tracking_mode = MONTE_MC_TRACK_CLEARANCE).
Not a mode of the reference (it tracks with the maximum over all materials, CBCT_real325im.cu:866-868); it samples
the same collision sites with fewer virtual collisions, so parity is (1) history by history against the oracle's
restatement of the same loop on the same Philox variates and (2) statistical against the reference's own loop
(GLOBAL mode).  The same bodies run on the CPU under SIMT emulation (tests/test_emu_mc.py).
"""
import math

import numpy as np
import pytest

import test_mc_gpu as G
from monte_b200 import _abi, scenes

pytestmark = pytest.mark.gpu


def _oracle_opts(monte, oracle, vol, lab, xs, seed, rng=None):
    grid, heavy = monte.clearance_grid(vol, lab, xs)
    o = oracle.mc_opts(oracle.RNG_PHILOX if rng is None else rng, seed=seed)
    return oracle.with_clearance(o, grid, heavy)


@pytest.mark.parametrize("cell_log2,poly,rayleigh", [(0, True, False), (1, True, False), (2, False, False), (1, True, True)])
def test_clearance_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly, rayleigh, mode=_abi.TRACK_CLEARANCE):
    g, vol, lab = G.scene(n=41, pitch=0.5, det=17, views=3, mode=_abi.SOURCE_CONE)
    vol.tracking_mode, vol.clearance_cell_log2 = mode, cell_log2
    xs = scenes.make_xs()
    if rayleigh:
        g.coherent_mode = _abi.COHERENT_FORMFACTOR
        scenes.add_formfactors(xs)
    spec, keep = scenes.kramers_spectrum() if poly else (scenes.mono_spectrum(45.0), None)
    per, seed, view = 24, 13, 1
    sc = monte.Scene(g, vol, lab, xs, spec)
    f_gpu, e_gpu = sc.fates(view, per, seed)
    sc.close()
    o, keepg = _oracle_opts(monte, oracle, vol, lab, xs, seed)
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, o, per,
                                            views=(view, view + 1), want_fates=True)
    assert ((f_gpu & 0xFF) != 0).all()
    same = f_gpu == f_cpu
    assert same.mean() > 0.999, "only %.4f of %d histories end identically" % (same.mean(), same.size)
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)
    for k in (1, 2, 3, 4):
        assert ((f_cpu & 0xFF) == k).any(), k


@pytest.mark.parametrize("cell_log2,poly", [(0, True), (1, False)])
def test_adaptive_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly):
    """tracking_mode ADAPTIVE: the light majorant only where the clearance exceeds the break-even distance"""
    test_clearance_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly, False, mode=_abi.TRACK_ADAPTIVE)


@pytest.mark.parametrize("cell_log2,poly,rayleigh", [(0, True, False), (1, False, False), (1, True, True)])
def test_directional_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly, rayleigh):
    """tracking_mode DIRECTIONAL: one clearance per octant of the flight direction, re-read after every deflection"""
    test_clearance_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly, rayleigh, mode=_abi.TRACK_DIRECTIONAL)


def test_adaptive_never_needs_more_steps_than_the_reference_loop(monte):
    g, vol, lab = G.scene(n=41, pitch=0.5, det=17, views=2)
    xs = scenes.make_xs()
    kr, keep = scenes.kramers_spectrum()
    for spec in (scenes.mono_spectrum(140.0), scenes.mono_spectrum(60.0), kr):
        steps = {}
        for mode in (_abi.TRACK_GLOBAL, _abi.TRACK_CLEARANCE, _abi.TRACK_ADAPTIVE, _abi.TRACK_DIRECTIONAL):
            vol.tracking_mode, vol.clearance_cell_log2 = mode, 0
            _, _, st = monte.simulate(g, vol, lab, xs, spec, 300, 17)
            steps[mode] = st["woodcock_steps"] / st["histories"]
        assert steps[_abi.TRACK_ADAPTIVE] <= 1.03 * steps[_abi.TRACK_GLOBAL], steps
        assert steps[_abi.TRACK_ADAPTIVE] <= 1.03 * steps[_abi.TRACK_CLEARANCE], steps
        assert steps[_abi.TRACK_DIRECTIONAL] <= 0.93 * steps[_abi.TRACK_ADAPTIVE], steps      # the directional grids pay everywhere


def test_clearance_counters_and_fewer_steps_than_the_reference_loop(monte, oracle):
    g, vol, lab = G.scene(n=41, pitch=0.5, det=17, views=2)
    xs = scenes.make_xs()
    spec, keep = scenes.kramers_spectrum()
    per, seed = 80, 21
    r0, r5, st_ref = monte.simulate(g, vol, lab, xs, spec, per, seed)              # the reference's single majorant
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 1
    c0, c5, st = monte.simulate(g, vol, lab, xs, spec, per, seed)
    o, keepg = _oracle_opts(monte, oracle, vol, lab, xs, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec, o, per)
    n = st["histories"]
    assert np.abs(c0.astype(int) - o0).sum() <= 0.001 * n and np.abs(c5.astype(int) - o5).sum() <= 0.001 * n
    for k in ("primaries", "scatter_detected", "absorbed", "interactions", "coherent", "compton", "woodcock_steps"):
        assert abs(st[k] - res[k]) <= 0.001 * max(res[k], 1) + 5, (k, st[k], res[k])
    # the point of the mode: far fewer tentative collisions for the same physics
    assert st["woodcock_steps"] < 0.6 * st_ref["woodcock_steps"], (st["woodcock_steps"], st_ref["woodcock_steps"])
    # same physics: totals agree statistically with the single-majorant run (independent variate use)
    for k in ("primaries", "scatter_detected", "absorbed", "compton", "coherent"):
        assert abs(st[k] - st_ref[k]) < 5 * math.sqrt(st[k] + st_ref[k] + 1), (k, st[k], st_ref[k])
    c2, dof, frac3 = G.chi2_images(c0, r0, binomial_per=per)
    assert abs(c2 - dof) < 5 * math.sqrt(2 * dof), ("image0 chi2", c2, dof)


def test_clearance_partition_and_label_update(monte):
    """photon ranges give the undivided tallies; re-uploading other labels rebuilds the grid"""
    g, vol, lab = G.scene(n=33, pitch=1.0, det=9, views=1)
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 1
    xs = scenes.make_xs()
    spec = scenes.mono_spectrum(50.0)
    a0, a5, st = monte.simulate(g, vol, lab, xs, spec, 40, 5)
    b0, b5, _ = monte.simulate(g, vol, lab, xs, spec, 40, 5, n_range=(0, 17))
    c0, c5, _ = monte.simulate(g, vol, lab, xs, spec, 40, 5, n_range=(17, 40))
    assert np.array_equal(a0, b0 + c0) and np.array_equal(a5, b5 + c5)
    # one material only: nothing to exclude, the mode silently is the reference's loop
    xs1 = scenes.make_xs(("h2o",))
    d0, d5, st1 = monte.simulate(g, vol, (lab > 0).astype(np.uint8), xs1, spec, 40, 5)
    vol.tracking_mode = _abi.TRACK_GLOBAL
    e0, e5, st2 = monte.simulate(g, vol, (lab > 0).astype(np.uint8), xs1, spec, 40, 5)
    assert np.array_equal(d0, e0) and np.array_equal(d5, e5) and st1["woodcock_steps"] == st2["woodcock_steps"]
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 11
    with pytest.raises(monte.MonteError, match="clearance_cell_log2"):
        monte.simulate(g, vol, lab, xs, spec, 2, 1)


# ---- device-resident projector (monte_gpu_project_primary_dev): BASELINE config 3 without the host round trip
def _dev_buffers(monte, shapes):
    """device float32 buffers for the *_dev entry points: CUDA tensors on the GPU, host arrays under emulation"""
    if hasattr(monte, "Dev"):
        arrs = [np.full(s, np.nan, np.float32) for s in shapes]
        return [monte.Dev(a) for a in arrs], (lambda d: d.a), (lambda: None)
    import torch
    ts = [torch.full(tuple(int(x) for x in s), float("nan"), dtype=torch.float32, device="cuda") for s in shapes]
    return ts, (lambda t: t.cpu().numpy()), torch.cuda.synchronize


def test_resident_projector_feeds_fdk_like_the_host_path(monte, oracle):
    """projection on the device -> filter -> backprojection equals project_primary (host buffers) -> fdk, bit for
    bit; the projection itself equals the oracle's within 1e-4 (north_star)"""
    lab = scenes.cylinder_phantom(33, 0.8)
    vol = scenes.volume_for(lab, 0.8)
    xs = scenes.make_xs()
    n_views, nu, nv = 24, 40, 24
    g = scenes.mc_geom(0, 32.5 / nu, n_views=n_views, ny=nu, nx=nv)              # ny = transaxial = FDK's iu
    fg = _abi.generic_fdk_geom(n_views, nu, nv, 32, textbook=True)
    fg.angle0_deg, fg.angle_step_deg = g.angle0_deg, g.angle_step_deg
    host_map = monte.project_primary(g, vol, lab, xs, 70.0)
    ref = oracle.project_primary(g, vol, lab, oracle.tables_from_xs(xs), 70.0)
    assert np.abs(host_map - ref).max() <= 1e-4 * np.abs(ref).max()
    _, vol_host, _, _ = monte.fdk(fg, host_map, want_filtered=False)
    (d_map, d_filt, d_vol), to_np, sync = _dev_buffers(monte, [(n_views, nu, nv), monte.fdk_filtered_shape(fg), (fg.nz, fg.ny, fg.nx)])
    pr = monte.Projector(vol, lab)
    pr.project(g, xs, 70.0, d_map, views=(0, 9))                                  # in two view ranges, as a sharded host would
    pr.project(g, xs, 70.0, d_map, views=(9, n_views))
    monte.fdk_filter_dev(fg, d_map, d_filt)
    monte.fdk_backproject_dev(fg, d_filt, d_vol)
    sync()
    pr.close()
    assert np.array_equal(to_np(d_map), host_map)
    assert np.array_equal(to_np(d_vol), vol_host)
    with pytest.raises(monte.MonteError, match="view range"):
        pr2 = monte.Projector(vol, lab)
        try:
            pr2.project(g, xs, 70.0, d_map, views=(3, 99))
        finally:
            pr2.close()
