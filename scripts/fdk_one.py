#!/usr/bin/env python
"""one C3 backprojection with the kernel MONTE_BP_VARIANT selects (for ncu captures)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api  # noqa: E402

api.init(0)
g = _abi.generic_fdk_geom(720, 1024, 768, 512)
proj = torch.rand((g.n_views, g.nu, g.nv), device="cuda")
filt = torch.zeros(api.fdk_filtered_shape(g), device="cuda")
vol = torch.empty((g.nz, g.ny, g.nx), device="cuda")
api.fdk_filter_dev(g, proj, filt)
for _ in range(int(os.environ.get("REPS", "2"))):
    api.fdk_backproject_dev(g, filt, vol)
torch.cuda.synchronize()
