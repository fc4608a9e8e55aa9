"""CPU, world_size 2, gloo: the multi-GPU partition logic of monte_b200/dist.py with the oracle
standing in for the kernels.  Sharded == unsharded, bit for bit (integer tallies; per-voxel FDK)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from monte_b200 import _abi, scenes  # noqa: E402
from monte_b200 import dist as mdist  # noqa: E402


def test_split_range_covers_everything():
    for n in (0, 1, 7, 360, 947, 1024):
        for ws in (1, 2, 3, 8):
            parts = [mdist.split_range(n, ws, r) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _mc_scene():
    lab = scenes.cylinder_phantom(17, 2.0)
    g = scenes.mc_geom(9, 32.5 / 9, n_views=2)
    g.angle_step_deg = 90.0
    return g, scenes.volume_for(lab, 2.0), lab


def _worker(rank, ws, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from oracle import binding as ob
    # ---------------- MC: photon ranges + one reduce
    g, vol, lab = _mc_scene()
    tb = ob.tables_from_xs(scenes.make_xs())
    per = 31
    im0 = torch.zeros((g.n_views, g.ny, g.nx), dtype=torch.int32)
    im5 = torch.zeros_like(im0)

    def run_local(a0, a5, per_, views, n_range):
        o0, o5, _, _, _ = ob.mc_run(g, vol, lab, tb, scenes.mono_spectrum(), ob.mc_opts(ob.RNG_PHILOX, seed=4, n_threads=1),
                                    per_, views=views, n_range=n_range)
        a0 += torch.from_numpy(o0)
        a5 += torch.from_numpy(o5)

    for v in range(g.n_views):
        nr = mdist.mc_sharded_step(run_local, im0, im5, per, (v, v + 1))
        assert nr == mdist.split_range(per, ws, rank)
    # ---------------- FDK: view-sharded filter, gather, z-slab backprojection
    fg = _abi.generic_fdk_geom(10, 17, 9, 12)
    proj = np.random.default_rng(3).random((fg.n_views, fg.nu, fg.nv), dtype=np.float32)
    pitch = ((fg.nu + 1 + 3) // 4) * 4
    rows = torch.zeros((fg.n_views * fg.nv + 2, pitch))
    slab_box = {}

    def filter_views(lo, hi):
        if hi > lo:
            f = ob.fdk_filter(fg, proj)           # oracle filters everything; keep only this rank's views
            rows[: fg.n_views * fg.nv].view(fg.n_views, fg.nv, pitch)[lo:hi, :, : fg.nu] = torch.from_numpy(f[lo:hi])

    def backproject(z_lo, z_hi):
        dense = rows[: fg.n_views * fg.nv].view(fg.n_views, fg.nv, pitch)[:, :, : fg.nu].contiguous().numpy()
        g2 = fg.copy()
        g2.z_begin, g2.z_end = z_lo, z_hi
        slab_box["v"] = ob.fdk_backproject(g2, dense)[z_lo:z_hi]

    vr, zr = mdist.fdk_sharded(filter_views, lambda: None, backproject, rows, fg.n_views, fg.nv, fg.nz)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), im0=im0.numpy(), im5=im5.numpy(), slab=slab_box["v"], z=np.array(zr), v=np.array(vr))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_sharded_equals_unsharded(tmp_path, oracle, ws):
    port = 29500 + os.getpid() % 2000 + ws
    mp.spawn(_worker, args=(ws, port, str(tmp_path)), nprocs=ws, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(ws)]
    # MC: rank 0 holds the reduced tallies == undivided run
    g, vol, lab = _mc_scene()
    tb = oracle.tables_from_xs(scenes.make_xs())
    o0, o5, _, _, _ = oracle.mc_run(g, vol, lab, tb, scenes.mono_spectrum(), oracle.mc_opts(oracle.RNG_PHILOX, seed=4), 31)
    assert np.array_equal(parts[0]["im0"], o0) and np.array_equal(parts[0]["im5"], o5)
    # FDK: slabs tile the volume and equal the undivided reconstruction
    fg = _abi.generic_fdk_geom(10, 17, 9, 12)
    proj = np.random.default_rng(3).random((fg.n_views, fg.nu, fg.nv), dtype=np.float32)
    _, vol_o, _ = oracle.fdk(fg, proj)
    z = 0
    for p in parts:
        assert p["z"][0] == z
        assert np.array_equal(p["slab"], vol_o[p["z"][0]:p["z"][1]])
        z = p["z"][1]
    assert z == fg.nz
    assert sum(int(p["v"][1] - p["v"][0]) for p in parts) == fg.n_views
