"""The reference's OWN CUDA program on this GPU as a parity pin (VERDICT r1, row J3 / "parity" hole 2): the primary tally
`image0` of monte_cu/CBCT_real325.cu (`projection` kernel, cuRAND MRG32k3a, analytic phantom; built unmodified but for
two #define literals by oracle/Makefile, see scripts/ref_cuda_run.py) is quirk-free physics -- unscattered photons per
pixel -- and is compared here with libmonte_gpu's image0 of the same phantom: chi-square of two independent binomial
samples, and both against the deterministic line integrals.  Skipped where oracle/_ref holds no CUDA build of the
reference (it is built in the development container, where /root/reference is mounted, and travels with the snapshot)."""
import importlib.util
import os
import tempfile

import numpy as np
import pytest

from monte_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _helpers():
    spec = importlib.util.spec_from_file_location("ref_cuda_run", os.path.join(ROOT, "scripts", "ref_cuda_run.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_primary_tally_of_the_reference_cuda_kernel(monte):
    R = _helpers()
    exe = os.path.join(R.REF, "CBCT_real325_p100")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/CBCT_real325_p100 is not built (make -C oracle ref_cuda needs /root/reference)")
    per, views = 100, R.REF_CUDA_VIEWS
    with tempfile.TemporaryDirectory() as d:
        R.write_inputs(d, scenes.cylinder_phantom(200, 0.1, radius=8.5))
        wall, r0, r5, out, rc = R.run_ref(exe, d, "teth%dcyu8e", views, 300)
    assert rc == 0 and r0 is not None and r0.shape == (views, 325, 325), out[-400:]
    assert 0 < r0[0].max() <= per and (r5 >= r0).all()
    # the same phantom (CBCT_real325.cu:916-921: cylinder along x, rods in the y-z plane) voxelised at 0.05 cm
    lab = np.ascontiguousarray(scenes.cylinder_phantom(400, 0.05).transpose(2, 1, 0))
    g = scenes.mc_geom(325, 0.1, n_views=views)
    g.angle_step_deg = 1.0
    vol = scenes.volume_for(lab, 0.05)
    xs = scenes.make_xs()
    o0, o5, st = monte.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=5, views=(0, 1))
    # view 0 (the only one where the reference's start point, 10.1 cm before the axis, lies outside this phantom), pixels
    # away from rays that graze the analytic surfaces (where a voxelised and an analytic phantom differ by centimetres)
    line = monte.project_primary(g, vol, lab, xs, 140.0, views=(0, 1))[0].astype(np.float64)
    lab10 = np.ascontiguousarray(scenes.cylinder_phantom(200, 0.1).transpose(2, 1, 0))
    line10 = monte.project_primary(g, scenes.volume_for(lab10, 0.1), lab10, xs, 140.0, views=(0, 1))[0].astype(np.float64)
    keep = np.abs(line - line10) < 0.002
    assert keep.mean() > 0.8
    # The reference's tally is not an ideal binomial sample (profiles/r02_ref_cuda_dispersion.json: against our run -- whose two
    # seeds agree with each other at chi2/dof = 1.00 -- it is under-dispersed where the transmission is below 0.1, 0.93, and
    # over-dispersed on lightly attenuated rays, 2.2, where its analytic surfaces and our voxels differ; its own 100- and
    # 300-photon builds differ by 2.7 % in the first band).  So the bar is a band around 1, not a z-score.
    chi2, dof, z = R.chi2_images(r0[0][keep], o0[0][keep], per)
    assert dof > 40000 and 0.85 < chi2 / dof < 1.15, (chi2, dof, z)
    pexp = np.exp(-line)
    sig = np.maximum(np.sqrt(per * pexp * (1 - pexp)), 0.5)
    for name, im in (("reference", r0[0]), ("ours", o0[0])):
        frac = float((np.abs(im - per * pexp) / sig <= 3)[keep].mean())
        assert frac > 0.99, (name, frac)
    tot_r, tot_o = int(r0[0][keep].sum()), int(o0[0][keep].sum())
    assert abs(tot_r - tot_o) < 5.0 * np.sqrt(tot_r + tot_o), (tot_r, tot_o)
