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
def test_clearance_history_coupled_fates_match_oracle(monte, oracle, cell_log2, poly, rayleigh):
    g, vol, lab = G.scene(n=41, pitch=0.5, det=17, views=3, mode=_abi.SOURCE_CONE)
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, cell_log2
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
    assert same.mean() > 0.995, "only %.4f of %d histories end identically" % (same.mean(), same.size)
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)
    for k in (1, 2, 3, 4):
        assert ((f_cpu & 0xFF) == k).any(), k


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
    assert np.abs(c0.astype(int) - o0).sum() <= 0.005 * n and np.abs(c5.astype(int) - o5).sum() <= 0.005 * n
    for k in ("primaries", "scatter_detected", "absorbed", "interactions", "coherent", "compton", "woodcock_steps"):
        assert abs(st[k] - res[k]) <= 0.005 * max(res[k], 1) + 5, (k, st[k], res[k])
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
