#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/fdk_dist_breakdown.py: per-rank CUDA-event times of the phases of one sharded C3
reconstruction in the one-process-per-GPU form (filter own views | pack | all_to_all of the row bands | unpack | pad |
backprojection of the own z ranges), to see where a step's time goes at each N."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api, dist as mdist  # noqa: E402


def main():
    rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    api.init(local)
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    proj = torch.rand((g.n_views, g.nu, g.nv), device=dev)
    z_ranges = mdist.fdk_z_partition(g, ws)
    my_z = z_ranges[rank]
    n_my = sum(b - a for a, b in my_z)
    filt = torch.zeros(api.fdk_filtered_shape(g), device=dev)
    slab = torch.empty((max(n_my, 1), g.ny, g.nx), device=dev)
    ev = {}

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.setdefault(name, []).append(e)

    # the body of dist.fdk_sharded_band with marks between its phases
    def step():
        pieces = [mdist.split_range(g.n_views, ws, r) for r in range(ws)]
        v_lo, v_hi = pieces[rank]
        mark("start")
        api.fdk_filter_dev(g, proj, filt, v_lo, v_hi, pad=False)
        mark("filtered")
        pitch = filt.shape[1]
        f3 = filt[: g.n_views * g.nv].view(g.n_views, g.nv, pitch)
        need = []
        for r in range(ws):
            rr = []
            for z_lo, z_hi in z_ranges[r]:
                a, b = api.fdk_slab_rows(g, z_lo, z_hi)
                if b > a:
                    rr += [(a, b + 1), (0, 4)]
            need.append(mdist._merge_rows(rr, g.nv))
        n_rows = [sum(b - a for a, b in need[r]) for r in range(ws)]
        mine = v_hi - v_lo
        in_splits = [mine * n_rows[r] * pitch if r != rank else 0 for r in range(ws)]
        out_splits = [(pieces[q][1] - pieces[q][0]) * n_rows[rank] * pitch if q != rank else 0 for q in range(ws)]
        parts = [f3[v_lo:v_hi, a:b, :].reshape(-1) for r in range(ws) if r != rank for a, b in need[r]]
        send = torch.cat(parts) if parts else filt.new_empty(0)
        recv = filt.new_empty(sum(out_splits))
        mark("packed")
        if ws > 1:
            dist.all_to_all_single(recv, send, out_splits, in_splits)
        mark("exchanged")
        o = 0
        for q in range(ws):
            if q == rank:
                continue
            nq = pieces[q][1] - pieces[q][0]
            for a, b in need[rank]:
                n = nq * (b - a) * pitch
                f3[pieces[q][0]:pieces[q][1], a:b, :] = recv[o:o + n].view(nq, b - a, pitch)
                o += n
        mark("unpacked")
        api.fdk_pad_dev(g, filt)
        off = 0
        for z_lo, z_hi in my_z:
            if z_hi > z_lo:
                api.fdk_backproject_dev(g, filt, slab[off:off + z_hi - z_lo], z_lo, z_hi)
                off += z_hi - z_lo
        mark("done")
        return sum(in_splits) * 4, sum(out_splits) * 4

    for it in range(5):
        dist.barrier()
        sb, rb = step()
    torch.cuda.synchronize()
    names = ["start", "filtered", "packed", "exchanged", "unpacked", "done"]
    out = {"rank": rank, "ws": ws, "z": my_z, "send_MB": sb / 1e6, "recv_MB": rb / 1e6}
    for a, b in zip(names, names[1:]):
        out[a + "->" + b] = round(min(ev[a][i].elapsed_time(ev[b][i]) for i in range(2, 5)), 3)
    out["total"] = round(min(ev["start"][i].elapsed_time(ev["done"][i]) for i in range(2, 5)), 3)
    gathered = [None] * ws
    dist.all_gather_object(gathered, out)
    if rank == 0:
        for o in gathered:
            print(json.dumps(o))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
