#!/usr/bin/env python
"""C3 through monte_gpu_fdk (pinned host buffers, one device) for several view-chunk counts of the host pipeline
(MONTE_FDK_CHUNKS): wall clock per reconstruction and the library's own breakdown; results must be bit-identical."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api  # noqa: E402

api.init(0)
g = _abi.generic_fdk_geom(720, 1024, 768, 512)
proj = torch.rand((g.n_views, g.nu, g.nv)).pin_memory()
vol = torch.empty((g.nz, g.ny, g.nx)).pin_memory()
ref = None
for chunks in [int(c) for c in (sys.argv[1:] or ["8", "12", "16", "8"])]:
    os.environ["MONTE_FDK_CHUNKS"] = str(chunks)
    api.fdk(g, proj.numpy(), want_filtered=False, out=vol.numpy())
    t0 = time.perf_counter()
    for _ in range(3):
        _, _, _, st = api.fdk(g, proj.numpy(), want_filtered=False, out=vol.numpy())
    ms = (time.perf_counter() - t0) / 3 * 1e3
    same = True if ref is None else bool(np.array_equal(ref, vol.numpy()))
    if ref is None:
        ref = vol.numpy().copy()
    print(json.dumps({"chunks": chunks, "ms_per_reconstruction": ms, "ms_filter_phase": st["ms_filter"], "ms_backproject_tail": st["ms_backproject"],
                      "ms_d2h_tail": st["ms_d2h"], "bit_equal_to_first": same}), flush=True)
