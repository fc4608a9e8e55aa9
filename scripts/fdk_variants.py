#!/usr/bin/env python
"""A/B of the backprojection kernels on one B200 (MONTE_BP_VARIANT): 0 = L1 gathers (fdk_backproject_kernel),
10 / 11 = footprint staged in shared memory by the bulk-copy engine (fdk_backproject_smem_kernel, 4 / 3 CTAs per SM).
Every variant must give the bits of variant 0.  Prints one JSON line per (geometry, variant)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api  # noqa: E402

api.init(0)
cases = [("ragged", _abi.generic_fdk_geom(31, 96, 40, 40)), ("small", _abi.generic_fdk_geom(48, 65, 65, 32)),
         ("c5_256", _abi.generic_fdk_geom(360, 384, 384, 256)), ("c3", _abi.generic_fdk_geom(720, 1024, 768, 512))]
if "--big" in sys.argv:
    cases.append(("c5_1024", _abi.generic_fdk_geom(1440, 1536, 1536, 1024)))
variants = [int(v) for v in os.environ.get("VARIANTS", "0,10,11").split(",")]
for name, g in cases:
    gen = torch.Generator(device="cuda").manual_seed(7)
    proj = torch.rand((g.n_views, g.nu, g.nv), device="cuda", generator=gen)
    filt = torch.zeros(api.fdk_filtered_shape(g), device="cuda")
    api.fdk_filter_dev(g, proj, filt)
    ref = None
    for var in variants:
        os.environ["MONTE_BP_VARIANT"] = str(var)
        for chunk in ([None] if var == 0 or "--chunks" not in sys.argv else [None, "0"]):
            if chunk is None:
                os.environ.pop("MONTE_BP_VCHUNK", None)
            else:
                os.environ["MONTE_BP_VCHUNK"] = chunk
            vol = torch.full((g.nz, g.ny, g.nx), float("nan"), device="cuda")
            api.fdk_backproject_dev(g, filt, vol)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                api.fdk_backproject_dev(g, filt, vol)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            if ref is None:
                ref = vol.clone()
            upd = g.nx * g.ny * g.nz * g.n_views
            print(json.dumps({"case": name, "variant": var, "vchunk": chunk, "ms": best, "gups": upd / best / 1e6,
                              "bit_equal_to_variant_0": bool(torch.equal(vol, ref)), "finite": bool(torch.isfinite(vol).all())}), flush=True)
    del proj, filt, ref, vol
    torch.cuda.empty_cache()
os.environ.pop("MONTE_BP_VARIANT", None)
