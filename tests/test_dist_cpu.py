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


# ---------------------------------------------------------------- pipelined exchange (async broadcasts)
def _fake_rows(v, nv, pitch, nu):
    rng = np.random.default_rng(1000 + v)
    r = np.zeros((nv, pitch), np.float32)
    r[:, :nu] = rng.random((nv, nu), dtype=np.float32)
    return r


def _pipe_worker(rank, ws, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    n_views, nv, nu, pitch, nz = 11, 5, 6, 8, 7
    rows = torch.full((n_views * nv + 2, pitch), float("nan"))        # NaN = "not received yet"
    rows[n_views * nv:] = 0
    slab = {}
    log = []

    def filter_views(lo, hi):
        for v in range(lo, hi):
            rows[v * nv:(v + 1) * nv] = torch.from_numpy(_fake_rows(v, nv, pitch, nu))

    def pad_views(lo, hi):                      # the dup column needs the NEXT row, i.e. the next piece's first row
        r1 = min(hi * nv + 2, n_views * nv + 2) if hi < n_views else n_views * nv + 2
        for r in range(lo * nv, r1):
            if r < n_views * nv:
                rows[r, nu] = rows[r + 1, 0] if r + 1 < n_views * nv else 0.0
                rows[r, nu + 1:] = 0

    def backproject_views(z_lo, z_hi, v_lo, v_hi, cont):
        acc = slab["v"] if cont else torch.zeros(z_hi - z_lo, dtype=torch.float32)
        for v in range(v_lo, v_hi):             # order-sensitive fp32 accumulation; reads one row past the view
            blk = rows[v * nv:(v + 1) * nv + 1, : nu + 1]
            assert not torch.isnan(blk).any(), "view %d used before it (or its successor) arrived" % v
            acc = acc * np.float32(1.0001) + blk.sum() * torch.arange(z_lo, z_hi, dtype=torch.float32)
        slab["v"] = acc
        log.append((v_lo, v_hi, cont))

    z_ranges = mdist.balanced_split([5, 1, 1, 1, 1, 1, 5], ws) if out_dir.endswith("uneven") else None
    vr, zr = mdist.fdk_sharded_pipelined(filter_views, pad_views, backproject_views, rows, n_views, nv, nz, z_ranges=z_ranges)
    np.savez(os.path.join(out_dir, "p%d.npz" % rank), slab=slab["v"].numpy(), z=np.array(zr), log=np.array(log))
    if out_dir.endswith("uneven"):               # several ranges per rank (balanced_blocks): each is its own slab
        multi = [[(0, 1), (5, 7)], [(1, 3)], [(3, 5)]]
        slabs = {}

        def bp_multi(z_lo, z_hi, v_lo, v_hi, cont):
            slab["v"] = slabs.get(z_lo)
            backproject_views(z_lo, z_hi, v_lo, v_hi, cont)
            slabs[z_lo] = slab["v"]
        rows[: n_views * nv] = float("nan")
        mdist.fdk_sharded_pipelined(filter_views, pad_views, bp_multi, rows, n_views, nv, nz, z_ranges=multi)
        np.savez(os.path.join(out_dir, "m%d.npz" % rank), **{"z%d_%d" % (a, b): slabs[a].numpy() for a, b in multi[rank]})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ws,uneven", [(1, False), (2, False), (3, False), (3, True)])
def test_pipelined_exchange_equals_sequential(tmp_path, ws, uneven):
    port = 29800 + os.getpid() % 1500 + ws + (7 if uneven else 0)
    if uneven:                                   # z-slabs of equal work instead of equal thickness
        tmp_path = tmp_path / "uneven"
        tmp_path.mkdir()
    mp.spawn(_pipe_worker, args=(ws, port, str(tmp_path)), nprocs=ws, join=True)
    n_views, nv, nu, pitch, nz = 11, 5, 6, 8, 7
    rows = np.zeros((n_views * nv + 2, pitch), np.float32)
    for v in range(n_views):
        rows[v * nv:(v + 1) * nv] = _fake_rows(v, nv, pitch, nu)
    for r in range(n_views * nv):
        rows[r, nu] = rows[r + 1, 0] if r + 1 < n_views * nv else 0.0
    acc = np.zeros(nz, np.float32)
    for v in range(n_views):
        acc = (torch.from_numpy(acc) * np.float32(1.0001) + torch.from_numpy(rows[v * nv:(v + 1) * nv + 1, : nu + 1]).sum()
               * torch.arange(0, nz, dtype=torch.float32)).numpy()
    z = 0
    for r in range(ws):
        p = np.load(os.path.join(str(tmp_path), "p%d.npz" % r))
        assert p["z"][0] == z
        assert np.array_equal(p["slab"], acc[p["z"][0]:p["z"][1]]), r
        assert [int(x) for x in p["log"][:, 0]] == [mdist.split_range(n_views, ws, q)[0] for q in range(ws)]
        z = p["z"][1]
    assert z == nz
    if uneven:
        seen = []
        for r in range(ws):
            m = np.load(os.path.join(str(tmp_path), "m%d.npz" % r))
            for k in m.files:
                a, b = (int(x) for x in k[1:].split("_"))
                assert np.array_equal(m[k], acc[a:b]), (r, k)
                seen += list(range(a, b))
        assert sorted(seen) == list(range(nz))


def test_balanced_split():
    """z-slabs of equal work: contiguous, complete, non-empty, and better balanced than equal thickness."""
    cost = [0.1] * 100 + [1.0] * 300 + [0.1] * 112                 # end slices nearly free, as at wide cone angles
    for ws in (1, 2, 3, 4, 8):
        parts = mdist.balanced_split(cost, ws)
        assert parts[0][0] == 0 and parts[-1][1] == len(cost)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(hi > lo for lo, hi in parts)
        load = [sum(cost[lo:hi]) for lo, hi in parts]
        even = [sum(cost[slice(*mdist.split_range(len(cost), ws, r))]) for r in range(ws)]
        assert max(load) <= max(even) + 1e-9
        assert max(load) <= 1.05 * sum(cost) / ws + 1.0
        aligned = mdist.balanced_split(cost, ws, align=16)
        assert aligned[0][0] == 0 and aligned[-1][1] == len(cost) and all(hi > lo for lo, hi in aligned)
        assert all(lo % 16 == 0 for lo, _ in aligned)
    assert mdist.balanced_split([1, 1, 1], 5)[-1][1] == 3          # fewer items than ranks: some pieces are empty
    assert [hi - lo for lo, hi in mdist.balanced_split([0, 0, 0, 0], 2)] == [1, 3]
    assert mdist.balanced_split([5, 0, 0, 0, 0, 0, 0, 5], 4) == [(0, 1), (1, 2), (2, 7), (7, 8)]


def test_fdk_slice_cost_matches_the_oracle_geometry():
    """the sampled on-detector fraction per slice agrees with a direct count using the projection
    formulas of recon/bp3d20.cpp:99-116"""
    from monte_b200 import _abi
    g = _abi.generic_fdk_geom(90, 96, 40, 48)
    c = mdist.fdk_slice_cost(g, overhead=0.0, partial_penalty=0.0, stride=1)
    X = (g.x0 + g.vox * np.arange(g.nx))[None, :, None]
    Y = (g.y0 - g.vox * np.arange(g.ny))[:, None, None]
    b = np.deg2rad(g.angle0_deg + g.angle_step_deg * np.arange(g.n_views))[None, None, :]
    k = g.dsd / (X * np.cos(b) + Y * np.sin(b) + g.dso)
    u_ok = np.abs(k * (-X * np.sin(b) + Y * np.cos(b))) <= g.half_u
    for z in (0, 5, 24, 47):
        Z = g.z0 - g.vox * z
        assert abs(c[z] - np.mean(u_ok & (np.abs(k * Z) <= g.half_v))) < 1e-12
    assert c[24] > c[0]                                            # central slices see the detector more often


def test_balanced_blocks_cover_every_slice_once_and_beat_the_contiguous_split():
    """cheap end blocks + expensive middle blocks (C3's shape): whole-block contiguous slabs cannot be
    balanced over 8 ranks, handing the cheap blocks out separately can"""
    cost = np.array([0.3] * 64 + [1.5] * 384 + [0.3] * 64)
    for ws in (2, 3, 4, 8):
        parts = mdist.balanced_blocks(cost, ws, 16)
        covered = sorted(z for pr in parts for a, b in pr for z in range(a, b))
        assert covered == list(range(512))
        assert all(a % 16 == 0 for pr in parts for a, _ in pr)
    load = lambda parts: max(sum(cost[a:b].sum() for a, b in pr) for pr in parts)
    contiguous = [[r] for r in mdist.balanced_split(cost, 8, 16)]
    assert load(mdist.balanced_blocks(cost, 8, 16)) < 0.9 * load(contiguous)
    assert mdist.balanced_blocks([1.0] * 40, 3, 16) == [[(0, 16)], [(16, 32)], [(32, 40)]]
    from monte_b200 import _abi
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    for ws in (1, 2, 4, 8):
        parts = mdist.fdk_z_partition(g, ws)
        assert sorted(z for pr in parts for a, b in pr for z in range(a, b)) == list(range(512))


# ---------------------------------------------------------------- band-limited exchange (one all_to_all)
def _band_rows(z_lo, z_hi):
    return (z_lo + 1, z_hi + 2)                  # a made-up "slab -> detector rows" map for the CPU test


def _band_slab_value(rows3, z_lo, z_hi, n_views, nv, nu):
    """order-sensitive fp32 sum over views of what the slab reads: the rows of its band with their duplicated
    column (= first element of the following row, hence the extra exchanged row) and rows 0..2 of the next
    view with theirs (zeros after the last view)"""
    a, b = _band_rows(z_lo, z_hi)
    acc = torch.zeros(z_hi - z_lo, dtype=torch.float32)
    for v in range(n_views):
        blk = rows3[v, a:b, : nu + 1]
        nxt = rows3[v + 1, 0:3, : nu + 1] if v + 1 < n_views else torch.zeros(3, nu + 1)
        assert not torch.isnan(blk).any() and not torch.isnan(nxt).any(), "view %d: a row the slab reads never arrived" % v
        acc = acc * np.float32(1.0001) + (blk.sum() + 2 * nxt.sum()) * torch.arange(z_lo, z_hi, dtype=torch.float32)
    return acc


def _band_worker(rank, ws, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    n_views, nv, nu, pitch = 11, 12, 6, 8
    rows = torch.full((n_views * nv + 2, pitch), float("nan"))
    rows[n_views * nv:] = 0
    z_ranges = [[(0, 2), (8, 9)], [(2, 5)], [(5, 8)]][:ws] if ws == 3 else [(0, 4), (4, 9)]
    out = {}

    def filter_views(lo, hi):
        for v in range(lo, hi):
            rows[v * nv:(v + 1) * nv] = torch.from_numpy(_fake_rows(v, nv, pitch, nu))

    def pad():                                   # NaN-tolerant: rows that never arrived stay NaN
        f = rows[: n_views * nv]
        f[:-1, nu] = f[1:, 0]
        f[-1, nu] = 0

    def backproject_slab(z_lo, z_hi):
        out[(z_lo, z_hi)] = _band_slab_value(rows[: n_views * nv].view(n_views, nv, pitch), z_lo, z_hi, n_views, nv, nu)

    mdist.fdk_sharded_band(filter_views, pad, backproject_slab, _band_rows, rows, n_views, nv, z_ranges)
    n_nan = int(torch.isnan(rows).sum())
    np.savez(os.path.join(out_dir, "b%d.npz" % rank), nan=n_nan, **{"z%d_%d" % k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_band_exchange_equals_full_exchange(tmp_path, ws):
    port = 29300 + os.getpid() % 1500 + ws
    mp.spawn(_band_worker, args=(ws, port, str(tmp_path)), nprocs=ws, join=True)
    n_views, nv, nu, pitch = 11, 12, 6, 8
    rows = np.zeros((n_views * nv, pitch), np.float32)
    for v in range(n_views):
        rows[v * nv:(v + 1) * nv] = _fake_rows(v, nv, pitch, nu)
    rows[:-1, nu] = rows[1:, 0]
    full = torch.from_numpy(rows).view(n_views, nv, pitch)
    seen, some_rows_skipped = [], False
    for r in range(ws):
        b = np.load(os.path.join(str(tmp_path), "b%d.npz" % r))
        some_rows_skipped |= int(b["nan"]) > 0
        for k in b.files:
            if k == "nan":
                continue
            z_lo, z_hi = (int(x) for x in k[1:].split("_"))
            assert np.array_equal(b[k], _band_slab_value(full, z_lo, z_hi, n_views, nv, nu).numpy()), (r, k)
            seen += list(range(z_lo, z_hi))
    assert sorted(seen) == list(range(9))
    assert some_rows_skipped                     # the point of the exercise: not every row travels
