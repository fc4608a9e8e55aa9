"""Diagnostic (one B200): dispersion of the primary tallies of the reference's CUDA kernel and of ours against each other and
against the deterministic transmission, at 100 and 300 photons per pixel."""
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("ref_cuda_run", os.path.join(ROOT, "scripts", "ref_cuda_run.py"))
R = importlib.util.module_from_spec(spec)
spec.loader.exec_module(R)
from monte_b200 import api, scenes  # noqa: E402


def main():
    api.init(0)
    views = R.REF_CUDA_VIEWS
    ref = {}
    with tempfile.TemporaryDirectory() as d:
        R.write_inputs(d, scenes.cylinder_phantom(200, 0.1, radius=8.5))
        for per in (100, 300):
            wall, r0, r5, out, rc = R.run_ref(os.path.join(R.REF, "CBCT_real325_p%d" % per), d, "teth%dcyu8e", views, 300)
            ref[per] = r0[0].copy()
    lab = np.ascontiguousarray(scenes.cylinder_phantom(400, 0.05).transpose(2, 1, 0))
    g = scenes.mc_geom(325, 0.1, n_views=views)
    g.angle_step_deg = 1.0
    vol = scenes.volume_for(lab, 0.05)
    xs = scenes.make_xs()
    line = api.project_primary(g, vol, lab, xs, 140.0, views=(0, 1))[0].astype(np.float64)
    lab10 = np.ascontiguousarray(scenes.cylinder_phantom(200, 0.1).transpose(2, 1, 0))
    line10 = api.project_primary(g, scenes.volume_for(lab10, 0.1), lab10, xs, 140.0, views=(0, 1))[0].astype(np.float64)
    keep = np.abs(line - line10) < 0.002
    pexp = np.exp(-line)
    out = {}
    for per in (100, 300):
        ours = [api.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), per, seed=s, views=(0, 1))[0][0] for s in (5, 6)]
        for band, m in (("all", keep), ("p<0.1", keep & (pexp < 0.1)), ("0.1<p<0.9", keep & (pexp > 0.1) & (pexp < 0.9)), ("p>0.9", keep & (pexp >= 0.9))):
            row = {"pixels": int(m.sum())}
            for name, a, b in (("ours_vs_ours", ours[0], ours[1]), ("ref_vs_ours", ref[per], ours[0])):
                c2, dof, z = R.chi2_images(a[m], b[m], per)
                row[name] = {"chi2_over_dof": c2 / max(dof, 1), "dof": dof, "z": z}
            for name, a in (("ref", ref[per]), ("ours", ours[0])):
                var = per * pexp[m] * (1 - pexp[m])
                ok = var > 1e-9
                row[name + "_dispersion_vs_projector"] = float((((a[m] - per * pexp[m]) ** 2)[ok] / var[ok]).mean())
                row[name + "_total"] = int(a[m].sum())
            out["per%d_%s" % (per, band)] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
