"""Time the FDK device kernels on one GPU (CUDA events, inputs resident): GUPS and filter GMAC/s."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--views", type=int, default=720)
    ap.add_argument("--nu", type=int, default=1024)
    ap.add_argument("--nv", type=int, default=768)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--slabs", type=int, default=0, help="time every z-slab of the N-device partition alone (backprojection only)")
    a = ap.parse_args()
    api.init(0)
    g = _abi.generic_fdk_geom(a.views, a.nu, a.nv, a.n)
    proj = torch.rand((a.views, a.nu, a.nv), device="cuda")
    filt = torch.empty(api.fdk_filtered_shape(g), device="cuda")
    vol = torch.empty((a.n, a.n, a.n), device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    res = []
    for it in range(a.iters + 1):
        ev[0].record()
        api.fdk_filter_dev(g, proj, filt)
        ev[1].record()
        api.fdk_backproject_dev(g, filt, vol)
        ev[2].record()
        torch.cuda.synchronize()
        if it:
            res.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
    if a.slabs:
        cuts = api.fdk_partition(g, a.slabs)
        out = {"slabs": a.slabs, "cuts": list(cuts), "whole_ms": min(r[1] for r in res), "ms": {}}
        for zs in ("1", "0"):
            os.environ["MONTE_BP_ZSTREAMS"] = zs
            per = []
            for i in range(a.slabs):
                z0, z1 = cuts[i], cuts[i + 1]
                best = 1e30
                for it in range(a.iters + 1):
                    ev[0].record()
                    api.fdk_backproject_dev(g, filt, vol[z0:z1], z0, z1)
                    ev[1].record()
                    torch.cuda.synchronize()
                    if it:
                        best = min(best, ev[0].elapsed_time(ev[1]))
                per.append(round(best, 3))
            out["ms"]["z_block_streams" if zs == "1" else "one_stream"] = per
        del os.environ["MONTE_BP_ZSTREAMS"]
        print(json.dumps(out))
    tf = min(r[0] for r in res)
    tb = min(r[1] for r in res)
    upd = a.n ** 3 * a.views
    print(json.dumps({"n": a.n, "views": a.views, "nu": a.nu, "nv": a.nv,
                      "filter_ms": tf, "backproject_ms": tb, "gups": upd / tb / 1e6,
                      "filter_gmacs": a.views * a.nv * a.nu * a.nu / tf / 1e6,
                      "vol_absmax": float(vol.abs().max())}))


if __name__ == "__main__":
    main()
