"""BASELINE config 5: FDK backprojection sweep 256^3..1024^3 x 360..1440 views, detector nu = nv = 1.5 N.
Prints one JSON line per case: GUPS (backprojection alone and filter+backprojection) and the fraction of
the FP32-issue roofline (35 lane-instr per update, 148 SM x 128 lanes x sm_max_mhz)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api  # noqa: E402

api.init(0)
peak = 148 * 128 * 1.965e9
for n in (256, 512, 1024):
    for views in (360, 720, 1440):
        nu = nv = int(1.5 * n)
        g = _abi.generic_fdk_geom(views, nu, nv, n)
        proj = torch.rand((views, nu, nv), device="cuda")
        filt = torch.empty(api.fdk_filtered_shape(g), device="cuda")
        vol = torch.empty((n, n, n), device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        best = None
        for it in range(3):
            ev[0].record()
            api.fdk_filter_dev(g, proj, filt)
            ev[1].record()
            api.fdk_backproject_dev(g, filt, vol)
            ev[2].record()
            torch.cuda.synchronize()
            t = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
            if it and (best is None or t[1] < best[1]):
                best = t
        upd = n ** 3 * views
        print(json.dumps({"n": n, "views": views, "nu": nu, "nv": nv, "filter_ms": round(best[0], 3), "backproject_ms": round(best[1], 3),
                          "gups_backproject": round(upd / best[1] / 1e6, 1), "gups_total": round(upd / (best[0] + best[1]) / 1e6, 1),
                          "fp32_issue_roofline_frac": round(upd / best[1] * 1e3 * 35 / peak, 3)}), flush=True)
        del proj, filt, vol
        torch.cuda.empty_cache()
