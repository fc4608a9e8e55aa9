"""GPU: the two hot paths chained the way the reference chains its programs
(MC / projector -> counts->map -> FDK), checked against physics, not against the oracle:
the TEXTBOOK-mode reconstruction of the water / calcium phantom must return the attenuation
coefficients of the cross-section tables (monte_cpp/xcom2.csv, Ca.csv at 140 keV)."""
import numpy as np
import pytest

from monte_b200 import _abi, scenes

pytestmark = pytest.mark.gpu


def fdk_geom_for(mg, n, vox):
    g = _abi.generic_fdk_geom(mg.n_views, mg.ny, mg.nx, n, textbook=True)
    g.du = g.dv = mg.pixel
    g.half_u = g.half_v = mg.half
    g.dso, g.dsd = mg.dso, mg.dso + mg.dod
    g.angle0_deg, g.angle_step_deg = mg.angle0_deg, mg.angle_step_deg
    g.vox = vox
    g.x0, g.y0, g.z0 = -0.5 * n * vox, 0.5 * n * vox, 0.5 * n * vox
    return g


def test_projector_to_fdk_recovers_mu(monte):
    lab = scenes.cylinder_phantom(129, 0.25)                  # 32 cm cube, water r=10 + Ca rods
    mg = scenes.mc_geom(129, 32.5 / 129, n_views=360)
    vol = scenes.volume_for(lab, 0.25)
    xs = scenes.make_xs()
    line = monte.project_primary(mg, vol, lab, xs, 140.0)     # [view][transaxial][axial] = the FDK input layout
    g = fdk_geom_for(mg, 96, 0.25)
    _, rec, _, _ = monte.fdk(g, line, want_filtered=False)
    mu_w = float(xs.total[0][140]) * 1.0
    mu_ca = float(xs.total[1][140]) * 1.55
    c = 48
    # water: a ring between the rods and the wall, central slices (cone-beam artefacts grow with |z|)
    zz, tt, ss = np.ogrid[:96, :96, :96]
    r = np.hypot((ss - c + 0.5) * 0.25, (tt - c + 0.5) * 0.25)
    mid = (np.abs(zz - c) < 8)
    water = mid & (r > 7.5) & (r < 9.0)
    assert abs(rec[np.broadcast_to(water, rec.shape)].mean() / mu_w - 1) < 0.03
    air = mid & (r > 10.8) & (r < 11.8)
    assert abs(rec[np.broadcast_to(air, rec.shape)].mean()) < 0.05 * mu_w
    # the rod on the +x axis (5 cm from the centre): its core must show calcium
    rod = mid & (np.hypot((ss - c + 0.5) * 0.25 - 5.0, (tt - c + 0.5) * 0.25) < 0.8)
    assert abs(rec[np.broadcast_to(rod, rec.shape)].mean() / mu_ca - 1) < 0.06


def test_mc_counts_to_map_to_fdk(monte):
    """noisy version of the same chain at BASELINE config 1 scale: MC primaries -> -ln(I/I0) -> FDK"""
    lab = scenes.cylinder_phantom(65, 0.5, rods=False)
    mg = scenes.mc_geom(65, 0.5, n_views=360)
    vol = scenes.volume_for(lab, 0.5)
    xs = scenes.make_xs()
    per = 4000
    im0, im5, st = monte.simulate(mg, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=1)
    assert st["histories"] == 360 * 65 * 65 * per
    line = monte.counts_to_map(im0, per)                       # CBCT_real325im.cu:267-285
    g = fdk_geom_for(mg, 64, 0.4)
    _, rec, _, _ = monte.fdk(g, line, want_filtered=False)
    mu_w = float(xs.total[0][140])
    zz, tt, ss = np.ogrid[:64, :64, :64]
    r = np.hypot((ss - 31.5) * 0.4, (tt - 31.5) * 0.4)
    core = (np.abs(zz - 32) < 6) & (r < 7.0)
    assert abs(rec[np.broadcast_to(core, rec.shape)].mean() / mu_w - 1) < 0.04
    # scatter adds counts: the scatter-contaminated map under-estimates the attenuation (cupping)
    line5 = monte.counts_to_map(im5, per)
    _, rec5, _, _ = monte.fdk(g, line5, want_filtered=False)
    assert rec5[np.broadcast_to(core, rec.shape)].mean() < rec[np.broadcast_to(core, rec.shape)].mean()
