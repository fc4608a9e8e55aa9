#!/usr/bin/env python
"""Run the reference's own CUDA programs (oracle/_ref/CBCT_real325im[_s], CBCT_real325[_s], built from the sources
under /root/reference by `make -C oracle ref_cuda`) on this box's GPU, next to libmonte_gpu on the equivalent scene.

    python scripts/ref_cuda_run.py [--variant im_s|s|im|plain] [--ncu] [--out gpurun_out/ref_cuda.json]

TEST / MEASUREMENT INFRASTRUCTURE ("faster than `projection` compiled for sm_100 on the same box").
Inputs the reference reads by hard-coded name are synthesized in a scratch directory from the packed tables
(monte_b200/data/xs_tables.npz; PMMA.txt = the water table, the repository ships no PMMA data):
  xcom2.txt / Ca.txt / PMMA.txt   201 tab-separated rows "coh com ab mua" (CBCT_real325im.cu:299-355)
  125kv_al2mm.txt / 125kv_al10mm.txt  251 zeros: no CDF bin matches, so every photon keeps the default 140 keV
                                   (:492-498) -- which is also what the shipped program effectively does (survey Q3)
  cyu8_2.raw                      200^3 uint8 labels @0.1 cm: water cylinder r = 8.5 cm with the 8 calcium rods (the
                                   program only looks labels up inside r <= 9 cm, :904); spher01.raw zeros (unused)
The `_s` binaries differ from the shipped source in two #define literals only (oracle/Makefile); their sample size is
read back from the build (REF_CUDA_PHOTONS / REF_CUDA_VIEWS below must match the Makefile defaults).
Outputs: histories/s of the whole program (wall clock, incl. file I/O: the reference has no timers), the `projection`
kernel's own duration when --ncu is given, the same scene through monte_gpu_simulate, and two parity figures that pin
the quirk-free primary physics to reference-produced output: chi^2 of image0 (ours vs theirs, independent RNGs) and
-ln(image0/per) of the reference against monte_gpu_project_primary.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
REF_CUDA_PHOTONS, REF_CUDA_VIEWS = 200, 36          # oracle/Makefile defaults of the `_s` builds
VARIANTS = {
    "im_s": ("CBCT_real325im_s", REF_CUDA_PHOTONS, REF_CUDA_VIEWS, "teth%dpmma8etim2"),
    "im": ("CBCT_real325im", 10000, 360, "teth%dpmma8etim2"),
    "s": ("CBCT_real325_s", REF_CUDA_PHOTONS, REF_CUDA_VIEWS, "teth%dcyu8e"),
    "plain": ("CBCT_real325", 10000, 360, "teth%dcyu8e"),
}


def write_inputs(d, lab):
    from monte_b200 import scenes
    h2o, ca = scenes.load_tables()
    for name, t in (("xcom2.txt", h2o), ("Ca.txt", ca), ("PMMA.txt", h2o)):
        with open(os.path.join(d, name), "w") as f:
            for k in range(201):
                f.write("%.9g\t%.9g\t%.9g\t%.9g\n" % (t[0, k], t[1, k], t[2, k], t[3, k]))
    for name in ("125kv_al2mm.txt", "125kv_al10mm.txt"):
        with open(os.path.join(d, name), "w") as f:
            f.write("\n".join(["0"] * 251) + "\n")
    lab.tofile(os.path.join(d, "cyu8_2.raw"))
    np.zeros(185 * 185 * 325, np.uint8).tofile(os.path.join(d, "spher01.raw"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="im_s", choices=sorted(VARIANTS))
    ap.add_argument("--ncu", action="store_true", help="second run under ncu for the kernel's own duration")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_cuda.json"))
    ap.add_argument("--timeout", type=int, default=900)
    args = ap.parse_args()
    exe_name, per, views, tag = VARIANTS[args.variant]
    exe = os.path.join(REF, exe_name)
    if not os.path.exists(exe):
        print(json.dumps({"unavailable": "%s not built (make -C oracle ref_cuda needs /root/reference)" % exe_name}))
        return
    from monte_b200 import _abi, api, scenes
    voxel = args.variant.startswith("im")
    lab = scenes.cylinder_phantom(200, 0.1, radius=8.5)
    npix = 325 * 325
    res = {"variant": args.variant, "binary": exe_name, "photons_per_pixel": per, "views": views, "histories": npix * per * views,
           "phantom": "200^3 labels @0.1 cm, water r=8.5 + 8 Ca rods" if voxel else "analytic cylinder r=10 along x + 8 Ca rods (in the kernel)"}
    with tempfile.TemporaryDirectory() as d:
        write_inputs(d, lab)
        t = time.perf_counter()
        p = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=args.timeout)
        wall = time.perf_counter() - t
        res.update(rc=p.returncode, wall_s=wall, whole_program_hist_per_s=res["histories"] / wall,
                   stdout_tail=p.stdout.decode(errors="replace")[-300:])
        f0 = os.path.join(d, "proj325_%s.raw" % (tag % 0))
        f5 = os.path.join(d, "proj325_%s.raw" % (tag % 5))
        r0 = np.fromfile(f0, np.int32).reshape(-1, 325, 325)[:views].copy() if os.path.exists(f0) else None
        r5 = np.fromfile(f5, np.int32).reshape(-1, 325, 325)[:views].copy() if os.path.exists(f5) else None
        if args.ncu:
            log = os.path.join(d, "ncu.csv")
            subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", log, exe],
                           cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=args.timeout)
            try:
                import csv
                for row in csv.DictReader(l for l in open(log) if not l.startswith("==")):
                    if "projection" in row.get("Kernel Name", ""):
                        v = float(row["Metric Value"].replace(",", ""))
                        unit = row.get("Metric Unit", "ns")
                        sec = v * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}.get(unit, 1e-9)
                        res.update(kernel_s=sec, kernel_hist_per_s=res["histories"] / sec)
            except Exception as ex:
                res["ncu_error"] = repr(ex)[:200]
    if r0 is None:
        res["error"] = "the reference wrote no projection file"
        print(json.dumps(res))
        return
    res["ref_primary_fraction"] = float(r0.sum()) / res["histories"]
    res["ref_image5_minus_image0_fraction"] = float(r5.sum() - r0.sum()) / res["histories"]
    if voxel:
        # the same scene through libmonte_gpu: 200^3 labels, corner-indexed at origin -10 (CBCT_real325im.cu:921)
        api.init(0)
        g = scenes.mc_geom(325, 0.1, n_views=views)
        g.angle_step_deg = 1.0                         # num_p counts whole degrees (:508)
        vol = scenes.volume_for(lab, 0.1)
        xs = scenes.make_xs(("h2o", "ca", "pmma"))
        sp = scenes.mono_spectrum(140.0)
        api.simulate(g, vol, lab, xs, sp, per, seed=1, views=(0, 1))      # warm-up (scene upload, module load)
        t = time.perf_counter()
        o0, o5, st = api.simulate(g, vol, lab, xs, sp, per, seed=1)
        ours_wall = time.perf_counter() - t
        res.update(ours_kernel_ms=st["ms_kernel"], ours_hist_per_s_kernel=st["histories"] / (st["ms_kernel"] * 1e-3),
                   ours_hist_per_s_call=st["histories"] / ours_wall, ours_primary_fraction=st["primaries"] / st["histories"],
                   ours_scatter_detected_fraction=st["scatter_detected"] / st["histories"])
        if "kernel_hist_per_s" in res:
            res["speedup_kernel"] = res["ours_hist_per_s_kernel"] / res["kernel_hist_per_s"]
        res["speedup_whole_program_vs_our_call"] = res["ours_hist_per_s_call"] / res["whole_program_hist_per_s"]
        # chi^2 of the two primary images (independent samples of the same binomial per pixel)
        a, b = r0.astype(np.float64), o0.astype(np.float64)
        m = (a + b) > 0
        # pixels that every photon reaches unattenuated (both == per) carry no variance: leave them out
        m &= ~((a == per) & (b == per))
        # pooled binomial variance: Var(a - b) = 2 per p (1 - p), p = (a + b) / (2 per)
        pp = (a[m] + b[m]) / (2.0 * per)
        var = 2.0 * per * pp * (1.0 - pp)
        ok = var > 0
        chi2 = float((((a[m] - b[m])[ok]) ** 2 / var[ok]).sum())
        dof = int(ok.sum())
        res.update(chi2_image0=chi2, chi2_dof=dof, chi2_z=(chi2 - dof) / np.sqrt(2.0 * dof) if dof else None)
        # deterministic check: -ln(image0 / per) of the REFERENCE against our line integrals, per-pixel 3 sigma
        line = api.project_primary(g, vol, lab, xs, 140.0)
        pexp = np.exp(-line.astype(np.float64))
        sig = np.sqrt(np.maximum(per * pexp * (1 - pexp), 1e-12))
        z = (a - per * pexp) / np.maximum(sig, 0.5)
        res.update(ref_vs_projector_frac_within_3sigma=float((np.abs(z) <= 3).mean()), ref_vs_projector_mean_z=float(z.mean()),
                   ours_vs_projector_frac_within_3sigma=float((np.abs((b - per * pexp) / np.maximum(sig, 0.5)) <= 3).mean()))
        api.shutdown()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
