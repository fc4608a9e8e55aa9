"""Time the MC transport kernel on one GPU (CUDA events, scene resident): histories/s."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--per", type=int, default=947)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cone", action="store_true")
    ap.add_argument("--poly", action="store_true")
    a = ap.parse_args()
    api.init(0)
    g, vol, lab = scenes.config_c2() if a.config == "c2" else scenes.config_c1()
    if a.cone:
        g.source_mode = _abi.SOURCE_CONE
    spec, keep = scenes.kramers_spectrum() if a.poly else (scenes.mono_spectrum(140.0), None)
    sc = api.Scene(g, vol, lab, scenes.make_xs(), spec)
    im0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32, device="cuda")
    im5 = torch.zeros_like(im0)
    stats = torch.zeros(16, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e30
    for it in range(a.iters + 1):
        stats.zero_()
        v0 = (it * a.views) % (g.n_views - a.views + 1)
        ev[0].record()
        sc.simulate_dev(im0, im5, a.per, seed=it + 1, views=(v0, v0 + a.views), d_stats=stats)
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            best = min(best, ev[0].elapsed_time(ev[1]))
    st = api.unpack_stats(stats.cpu().numpy().astype(np.uint64))
    n = st["histories"]
    print(json.dumps({"config": a.config, "per": a.per, "views": a.views, "histories": n, "ms": best,
                      "hist_per_s": n / best * 1e3, "steps_per_hist": st["woodcock_steps"] / n,
                      "interactions_per_hist": st["interactions"] / n,
                      "primary_frac": st["primaries"] / n, "scatter_det_frac": st["scatter_detected"] / n}))


if __name__ == "__main__":
    main()
