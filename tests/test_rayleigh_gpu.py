"""GPU parity tests of the optional Rayleigh form-factor deflection (monte_mc_geom.coherent_mode =
MONTE_MC_COHERENT_FORMFACTOR).  The reference has no such mode (its coherent event keeps the direction,
CBCT_real325im.cu:656-695), so parity here is against the oracle's restatement of the same sampler on the same
Philox variates, history by history, and against the closed-form angular distribution (tests/test_oracle_mc.py).
The default mode (FORWARD) is what every other MC test covers; the same bodies run on the CPU under SIMT
emulation in tests/test_emu_mc.py.
"""
import numpy as np
import pytest

import test_mc_gpu as G
from monte_b200 import _abi, scenes

pytestmark = pytest.mark.gpu


def _scene(keV_views=3):
    g, vol, lab = G.scene(n=33, pitch=1.0, det=17, views=keV_views)
    g.coherent_mode = _abi.COHERENT_FORMFACTOR
    xs = scenes.add_formfactors(scenes.make_xs())
    return g, vol, lab, xs


@pytest.mark.parametrize("keV,poly", [(60.0, False), (30.0, False), (0.0, True)])
def test_formfactor_history_coupled_fates_match_oracle(monte, oracle, keV, poly):
    g, vol, lab, xs = _scene()
    spec, keep = scenes.kramers_spectrum() if poly else (scenes.mono_spectrum(keV), None)
    per, seed, view = 24, 91, 2
    sc = monte.Scene(g, vol, lab, xs, spec)
    f_gpu, e_gpu = sc.fates(view, per, seed)
    sc.close()
    _, _, res, f_cpu, e_cpu = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec,
                                            oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per,
                                            views=(view, view + 1), want_fates=True)
    assert ((f_gpu & 0xFF) != 0).all()
    same = f_gpu == f_cpu
    assert same.mean() > 0.999, "only %.4f of %d histories end identically" % (same.mean(), same.size)
    assert np.allclose(e_gpu[same], e_cpu[same], rtol=2e-5)
    assert res["coherent"] > 0.02 * res["interactions"]            # the branch under test is exercised


def test_formfactor_images_counters_and_difference_from_forward_mode(monte, oracle):
    g, vol, lab, xs = _scene(keV_views=2)
    spec = scenes.mono_spectrum(50.0)
    per, seed = 60, 6
    im0, im5, st = monte.simulate(g, vol, lab, xs, spec, per, seed)
    o0, o5, res, _, _ = oracle.mc_run(g, vol, lab, oracle.tables_from_xs(xs), spec,
                                      oracle.mc_opts(oracle.RNG_PHILOX, seed=seed), per)
    n = st["histories"]
    assert np.array_equal(im0, o0) or np.abs(im0.astype(int) - o0).sum() <= 0.001 * n
    assert np.abs(im5.astype(int) - o5).sum() <= 0.001 * n
    for k in ("primaries", "scatter_detected", "absorbed", "interactions", "coherent", "compton", "woodcock_steps"):
        assert abs(st[k] - res[k]) <= 0.001 * max(res[k], 1) + 5, (k, st[k], res[k])
    # the same histories with the reference's undeflected coherent event: same primaries (they never interact),
    # different scatter
    g.coherent_mode = _abi.COHERENT_FORWARD
    f0, f5, stf = monte.simulate(g, vol, lab, xs, spec, per, seed)
    assert np.array_equal(f0, im0) and stf["primaries"] == st["primaries"]
    assert not np.array_equal(f5, im5)
    # coherent scattering is elastic: with every Compton event switched off ... not possible through the ABI;
    # instead: the energy carried by detected scatter never exceeds the source energy
    assert st["sum_e_scatter"] <= 50.0 * st["scatter_detected"] * (1 + 1e-6)


def test_formfactor_mode_needs_tables(monte):
    g, vol, lab, _ = _scene(keV_views=1)
    with pytest.raises(monte.MonteError, match="form-factor tables"):
        monte.simulate(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum(60.0), 2, 1)
    g.coherent_mode = 5
    with pytest.raises(monte.MonteError, match="coherent_mode"):
        monte.simulate(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum(60.0), 2, 1)
