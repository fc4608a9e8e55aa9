"""Time the backprojection of z-slabs of C3 on one GPU: calibration of monte_b200.dist.fdk_slice_cost and a
check of the equal-work partitions it produces for 2, 4 and 8 ranks (per-rank ms, max/mean)."""
import os, sys, json, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api, dist as mdist
api.init(0)
g = _abi.generic_fdk_geom(720, 1024, 768, 512)
proj = torch.rand((720, 1024, 768), device="cuda")
filt = torch.zeros(api.fdk_filtered_shape(g), device="cuda")
api.fdk_filter_dev(g, proj, filt)
c0 = mdist.fdk_slice_cost(g, overhead=0.0, partial_penalty=0.0)


def slab_ms(a, b):
    slab = torch.empty((b - a, 512, 512), device="cuda")
    ts = []
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); api.fdk_backproject_dev(g, filt, slab, a, b); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


slab_ms(0, 16)
for (a, b) in [(0, 64), (64, 128), (112, 160), (208, 256), (256, 304), (0, 512)]:
    print(json.dumps({"z": [a, b], "ms": slab_ms(a, b), "on_detector_slices": float(c0[a:b].sum()), "slices": b - a}))
for ws in (2, 4, 8):
    for name, parts in (("equal thickness", [[mdist.split_range(512, ws, r)] for r in range(ws)]),
                        ("equal work, contiguous", [[r] for r in mdist.balanced_split(mdist.fdk_slice_cost(g), ws, 16)]),
                        ("equal work, blocks", mdist.balanced_blocks(mdist.fdk_slice_cost(g), ws, 16))):
        ms = [sum(slab_ms(a, b) for a, b in pr) for pr in parts]
        print(json.dumps({"ranks": ws, "partition": name, "z": [[list(q) for q in pr] for pr in parts], "ms": [round(m, 2) for m in ms],
                          "max_over_mean": round(max(ms) / (sum(ms) / ws), 3)}))
