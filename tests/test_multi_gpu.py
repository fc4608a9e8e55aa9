"""Multi-device mode of the C ABI (monte_gpu_init(ndev > 1, ids)): the host-buffer calls shard over
the bound devices inside libmonte_gpu and must return the bits of the one-device call.

The test bodies take the binding as an argument: here they run on real B200s (skipped on a box with fewer GPUs
than they need); tests/test_emu_multi.py runs the same bodies on the CPU emulation with MONTE_EMU_DEVICES."""
import os
import subprocess

import numpy as np
import pytest

from monte_b200 import _abi, scenes


def _need(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs" % n)


def rebind(api, devs):
    api.init(devs if len(devs) > 1 else devs[0])
    assert api.load().monte_gpu_device_count() == len(devs)


FDK_CASES = [
    # n_views, nu, nv, n, roi / mask tweak
    (24, 40, 30, 48, None),
    (31, 65, 33, 40, "roi"),          # ragged view split, partial ROI + sphere mask
    (16, 300, 20, 32, None),          # FFT filter path (nu > 256)
    (12, 33, 65, 24, "tall"),         # tall detector: every slice is visible, equal-cost slabs
]


def body_fdk_multi_equals_single(api, n_dev, case):
    n_views, nu, nv, n, tweak = case
    g = _abi.generic_fdk_geom(n_views, nu, nv, n)
    if tweak == "roi":
        g.s_begin, g.s_end, g.t_begin, g.t_end, g.z_begin, g.z_end = 3, n - 5, 2, n - 1, 4, n - 6
        g.mask_cs = g.mask_ct = g.mask_cz = n // 2
        g.mask_r2 = (n // 2 - 2) ** 2
    proj = np.random.default_rng(n_views).random((n_views, nu, nv), dtype=np.float32)
    rebind(api, [0])
    f1, v1, z1, _ = api.fdk(g, proj, want_zy=True)
    rebind(api, list(range(n_dev)))
    try:
        f2, v2, z2, st = api.fdk(g, proj, want_zy=True)
        f3, v3, _, _ = api.fdk(g, proj, want_filtered=False)          # second call: cached partition, reused buffers
    finally:
        rebind(api, [0])
    assert np.array_equal(f1, f2), "filtered views differ"
    assert np.array_equal(v1, v2) and np.array_equal(v1, v3), "volume differs from the one-device call"
    assert np.array_equal(z1, z2), "transposed volume differs"
    assert f3 is None and st["voxel_updates"] == (g.s_end - g.s_begin) * (g.t_end - g.t_begin) * (g.z_end - g.z_begin) * n_views
    cuts = api.fdk_partition(g, n_dev)
    assert cuts[0] == 0 and cuts[-1] == g.nz and all(b >= a for a, b in zip(cuts, cuts[1:]))


def body_fdk_multi_chunked(api, n_dev, chunks, case=(37, 40, 30, 48, None)):
    """the chunk-interleaved pipeline of the multi-device reconstruction: `chunks` chunks of views per device (the
    library takes 4 when there are >= 16 views per device; forced here, so that ragged and empty chunks occur),
    every chunk gathered from its owner and backprojected as soon as it and its successor are filtered"""
    old = os.environ.get("MONTE_FDK_MULTI_CHUNKS")
    os.environ["MONTE_FDK_MULTI_CHUNKS"] = str(chunks)
    try:
        body_fdk_multi_equals_single(api, n_dev, case)
    finally:
        if old is None:
            del os.environ["MONTE_FDK_MULTI_CHUNKS"]
        else:
            os.environ["MONTE_FDK_MULTI_CHUNKS"] = old


def body_mc_multi_equals_single(api, n_dev, reduce_mode, per=37):
    lab = scenes.cylinder_phantom(33, 1.0)
    mg = scenes.mc_geom(9, 32.5 / 9, n_views=3)
    mvol = scenes.volume_for(lab, 1.0)
    xs = scenes.make_xs()
    sp = scenes.mono_spectrum(140.0)
    rebind(api, [0])
    a0, a5, sa, ma0, ma5 = api.simulate(mg, mvol, lab, xs, sp, per, 7, maps=True)
    assert np.array_equal(ma0, api.counts_to_map(a0, per)) and np.array_equal(ma5, api.counts_to_map(a5, per))
    old = os.environ.get("MONTE_MC_REDUCE")
    os.environ["MONTE_MC_REDUCE"] = reduce_mode
    rebind(api, list(range(n_dev)))
    try:
        b0, b5, sb, mb0, mb5 = api.simulate(mg, mvol, lab, xs, sp, per, 7, maps=True)
        c0, c5, sc = api.simulate(mg, mvol, lab, xs, sp, per, 7, views=(1, 3), n_range=(5, per - 3))   # ragged sub-range, no maps
    finally:
        rebind(api, [0])
        if old is None:
            del os.environ["MONTE_MC_REDUCE"]
        else:
            os.environ["MONTE_MC_REDUCE"] = old
    d0, d5, sd = api.simulate(mg, mvol, lab, xs, sp, per, 7, views=(1, 3), n_range=(5, per - 3))
    assert np.array_equal(a0, b0) and np.array_equal(a5, b5), "tallies differ from the one-device run"
    assert np.array_equal(ma0, mb0) and np.array_equal(ma5, mb5), "maps differ"
    for k in ("histories", "primaries", "scatter_detected", "absorbed", "interactions", "woodcock_steps"):
        assert sa[k] == sb[k], k
    assert np.array_equal(c0, d0) and np.array_equal(c5, d5) and sc["histories"] == sd["histories"] == 2 * 81 * (per - 8)
    assert not c0[0].any()                                            # views outside the request stay untouched


def body_label_cache(api):
    """the label volume is re-uploaded only when its bytes changed; a changed buffer (same address) must be seen"""
    lab = scenes.cylinder_phantom(33, 1.0).copy()
    mg = scenes.mc_geom(9, 32.5 / 9, n_views=1)
    mvol = scenes.volume_for(lab, 1.0)
    xs, sp = scenes.make_xs(), scenes.mono_spectrum(140.0)
    os.environ["MONTE_MC_LABEL_CACHE"] = "1"                          # (the default hashes only where a clearance grid or a presence scan needs it)
    try:
        a0, a5, sa = api.simulate(mg, mvol, lab, xs, sp, 60, 3)
        b0, b5, sb = api.simulate(mg, mvol, lab, xs, sp, 60, 3)        # cached labels
        assert np.array_equal(a0, b0) and np.array_equal(a5, b5)
        lab[:] = 0                                                    # all air, in place
        c0, _, st = api.simulate(mg, mvol, lab, xs, sp, 60, 3)
        assert (c0 == 60).all() and st["primaries"] == st["histories"]
    finally:
        del os.environ["MONTE_MC_LABEL_CACHE"]
    lab[:] = scenes.cylinder_phantom(33, 1.0)                         # default policy: the change in place is seen as well
    e0, e5, _ = api.simulate(mg, mvol, lab, xs, sp, 60, 3)
    assert np.array_equal(a0, e0) and np.array_equal(a5, e5)
    lab[:] = 0
    os.environ["MONTE_MC_LABEL_CACHE"] = "0"
    try:
        lab[:] = scenes.cylinder_phantom(33, 1.0)
        d0, d5, _ = api.simulate(mg, mvol, lab, xs, sp, 60, 3)
    finally:
        del os.environ["MONTE_MC_LABEL_CACHE"]
    assert np.array_equal(a0, d0) and np.array_equal(a5, d5)


def body_scattered_label_upload(api, n_dev):
    """several devices: every device takes 1/N of the label volume from the host and the rest from its peers over
    NVLink (label_gather_kernel).  Same tallies with the scatter on or off, with the label cache on or off, when the
    buffer changes in place, and when only some devices already hold the labels (a one-device call came first)."""
    lab = scenes.cylinder_phantom(33, 1.0).copy()
    lab[::3, 1::2, ::5] = 2                                            # make every part of the volume matter
    mg = scenes.mc_geom(9, 32.5 / 9, n_views=2)
    mvol = scenes.volume_for(lab, 1.0, tight=False)
    xs, sp = scenes.make_xs(), scenes.mono_spectrum(60.0)
    per = 40
    env = os.environ
    keep = {k: env.get(k) for k in ("MONTE_MC_LABEL_CACHE", "MONTE_MC_LABEL_SCATTER")}
    try:
        rebind(api, [0])
        env["MONTE_MC_LABEL_CACHE"] = "0"
        want = api.simulate(mg, mvol, lab, xs, sp, per, 11)[:2]
        rebind(api, list(range(n_dev)))
        for cache in ("0", "1"):
            for scatter in ("1", "0"):
                env["MONTE_MC_LABEL_CACHE"], env["MONTE_MC_LABEL_SCATTER"] = cache, scatter
                for rep in range(2):                                  # second call: cached (cache = 1) or re-scattered
                    got = api.simulate(mg, mvol, lab, xs, sp, per, 11)[:2]
                    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (cache, scatter, rep)
        # the buffer changes in place: a scattered re-upload must reach every device
        env["MONTE_MC_LABEL_CACHE"], env["MONTE_MC_LABEL_SCATTER"] = "1", "1"
        lab2 = lab.copy()
        lab2[lab2 == 2] = 1
        rebind(api, [0])
        want2 = api.simulate(mg, mvol, lab2, xs, sp, per, 11)[:2]      # device 0 now holds lab2 (hash known)
        assert not np.array_equal(want2[1], want[1])
        rebind(api, list(range(n_dev)))                                # rebinding resets the per-device scenes ...
        got2 = api.simulate(mg, mvol, lab2, xs, sp, per, 11)[:2]
        assert np.array_equal(got2[0], want2[0]) and np.array_equal(got2[1], want2[1])
        lab2[:] = lab                                                  # ... and now in place, same address
        got3 = api.simulate(mg, mvol, lab2, xs, sp, per, 11)[:2]
        assert np.array_equal(got3[0], want[0]) and np.array_equal(got3[1], want[1])
    finally:
        for k, v in keep.items():
            if v is None:
                env.pop(k, None)
            else:
                env[k] = v
        rebind(api, [0])


def body_hu_volume_multi_equals_single(api, n_dev):
    """a segmented CT volume tracked with the majorant of the classes present (one presence scan for all devices, scattered
    label upload) and the ring detector, sharded over the devices: the bits of the one-device runs"""
    hu = scenes.hu_head_phantom(33, 0.6)
    lab, xs, _, present = api.ctnum_segment(hu, api.hu_classes_default(True), scenes.make_xs(), 60.0)
    vol = scenes.volume_for(lab, 0.6)
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    sp = scenes.mono_spectrum(60.0)
    flat = scenes.mc_geom(9, 32.5 / 9, n_views=2)
    ring = scenes.ring_geom(12, 3, 4.0, 16.0, n_views=2)
    rebind(api, [0])
    want = [api.simulate(g, vol, lab, xs, sp, 23, 5)[:2] for g in (flat, ring)]
    rebind(api, list(range(n_dev)))
    try:
        got = [api.simulate(g, vol, lab, xs, sp, 23, 5)[:2] for g in (flat, ring)]
    finally:
        rebind(api, [0])
    for w, g_ in zip(want, got):
        assert np.array_equal(w[0], g_[0]) and np.array_equal(w[1], g_[1])
    assert want[1][1].sum() > want[1][0].sum() > 0


def body_argument_errors(api):
    from monte_b200.api import MonteError
    lab = scenes.cylinder_phantom(17, 2.0)
    mg = scenes.mc_geom(5, 6.5, n_views=2)
    mvol = scenes.volume_for(lab, 2.0)
    xs = scenes.make_xs()
    err = getattr(api, "MonteError", MonteError)
    for bad in (scenes.mono_spectrum(0.0), scenes.mono_spectrum(-5.0), scenes.mono_spectrum(250.0)):
        with pytest.raises(err, match="mono_keV"):
            api.simulate(mg, mvol, lab, xs, bad, 4, 1)
    xz = scenes.make_xs()
    for m in range(xz.n_materials):
        xz.total[m][60] = 0.0                                          # zero majorant at a reachable energy: must not spin
    with pytest.raises(err, match="majorant"):
        api.simulate(mg, mvol, lab, xz, scenes.mono_spectrum(140.0), 4, 1)
    api.simulate(mg, mvol, lab, xz, scenes.mono_spectrum(50.0), 4, 1)  # ... but 60 keV is out of reach from 50 keV
    # [v, v) is empty: nothing runs, nothing is written
    im0 = np.full((2, 5, 5), -7, np.int32)
    im5 = np.full((2, 5, 5), -7, np.int32)
    _, _, st = api.simulate(mg, mvol, lab, xs, scenes.mono_spectrum(140.0), 4, 1, views=(0, 0), out=(im0, im5))
    assert (im0 == -7).all() and (im5 == -7).all() and st["histories"] == 0
    with pytest.raises(err, match="ndev"):
        api.init(list(range(9)))
    api.init(0)


# ------------------------------------------------------------------------------------------ on real GPUs
@pytest.mark.gpu
@pytest.mark.parametrize("case", FDK_CASES[:3])
def test_fdk_two_devices_equal_one(monte, case):
    _need(2)
    body_fdk_multi_equals_single(monte, 2, case)


@pytest.mark.gpu
@pytest.mark.parametrize("chunks,case", [(8, (70, 40, 30, 48, None)), (4, (37, 40, 30, 48, None)), (3, (5, 24, 16, 16, None)), (4, (64, 300, 24, 32, None)), (2, (31, 65, 33, 40, "roi"))])
def test_fdk_two_devices_chunked_pipeline(monte, chunks, case):
    _need(2)
    body_fdk_multi_chunked(monte, 2, chunks, case)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_mc_two_devices_equal_one(monte, mode):
    _need(2)
    body_mc_multi_equals_single(monte, 2, mode)


@pytest.mark.gpu
def test_label_cache(monte):
    body_label_cache(monte)


@pytest.mark.gpu
def test_hu_volume_and_ring_detector_two_devices(monte):
    _need(2)
    body_hu_volume_multi_equals_single(monte, 2)


@pytest.mark.gpu
def test_scattered_label_upload_two_devices(monte):
    _need(2)
    body_scattered_label_upload(monte, 2)


@pytest.mark.gpu
def test_argument_errors(monte):
    body_argument_errors(monte)


@pytest.mark.gpu
def test_overlapping_launches_of_one_scene(monte):
    """two launches of one scene on two streams take their own work counters: the sum equals the serial run"""
    import torch
    lab = scenes.cylinder_phantom(65, 0.5)
    mg = scenes.mc_geom(65, 0.5, n_views=2)
    mvol = scenes.volume_for(lab, 0.5)
    sc = monte.Scene(mg, mvol, lab, scenes.make_xs(), scenes.mono_spectrum(140.0))
    per = 400
    ref0 = torch.zeros((2, 65, 65), dtype=torch.int32, device="cuda"); ref5 = torch.zeros_like(ref0)
    sc.simulate_dev(ref0, ref5, per, seed=5)
    torch.cuda.synchronize()
    for _ in range(3):
        a0 = torch.zeros_like(ref0); a5 = torch.zeros_like(ref0)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        sc.simulate_dev(a0, a5, per, seed=5, views=(0, 1), stream=s1)
        sc.simulate_dev(a0, a5, per, seed=5, views=(1, 2), stream=s2)
        torch.cuda.synchronize()
        assert torch.equal(a0, ref0) and torch.equal(a5, ref5)
    sc.close()


@pytest.mark.gpu
def test_drivers_with_two_gpus_write_the_one_gpu_bytes(tmp_path):
    """cbct_mc / cbct_fdk --gpus 2 (the C++ hosts, monte_gpu_init(2, NULL)) against --gpus 1, byte for byte"""
    _need(2)
    from monte_b200 import build
    from test_drivers import _write_csv
    build.build_lib()
    bins = {os.path.basename(p): p for p in build.build_drivers()}
    d = str(tmp_path)
    h2o, ca = scenes.load_tables()
    _write_csv(os.path.join(d, "xcom2.csv"), h2o)
    _write_csv(os.path.join(d, "Ca.csv"), ca)
    subprocess.check_call([bins["make_fantom"], "cylinder", "33", "1.0", os.path.join(d, "cyl.raw")])
    for tag, gpus in (("a", "1"), ("b", "2")):
        subprocess.run([bins["cbct_mc"], "cyl.raw", "33", "1.0", "xcom2.csv", "Ca.csv", "17", str(32.5 / 17), "4", "201", "3", tag,
                        "--gpus", gpus], cwd=d, check=True, stdout=subprocess.PIPE)
    for f in ("proj_%s0.raw", "proj_%s5.raw", "map_%s0.raw", "map_%s5.raw"):
        assert open(os.path.join(d, f % "a"), "rb").read() == open(os.path.join(d, f % "b"), "rb").read(), f
    np.random.default_rng(3).random((360, 65, 65), dtype=np.float32).tofile(os.path.join(d, "m.raw"))
    for tag, gpus in (("a", "1"), ("b", "2")):
        subprocess.run([bins["cbct_fdk"], "bp3d20", "m.raw", tag, "--gpus", gpus], cwd=d, check=True, stdout=subprocess.PIPE)
    for f in ("xy_%s.raw", "zy_%s.raw", "map_%s.raw"):
        assert open(os.path.join(d, f % "a"), "rb").read() == open(os.path.join(d, f % "b"), "rb").read(), f
