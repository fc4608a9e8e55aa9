"""GPU parity tests of the FDK path, called through the C ABI (libmonte_gpu.so).

Tolerance (BASELINE.json north_star): the FDK volume and the filtered projections agree with the
reference CPU path within 1e-4 relative, fp32.  "Relative" is to max|reference| of the compared
array (the volume crosses zero), and it must hold for EVERY element:
    max|gpu - ref| <= REL * max|ref|.
"""
import os

import numpy as np
import pytest

from monte_b200 import _abi
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
REL = 1e-4


def rand(seed, shape):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def assert_close(got, ref, what):
    scale = float(np.abs(ref).max())
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
    assert err <= REL * scale, "%s: max err %.3e > %.1e * max|ref| (%.3e)" % (what, err, REL, scale)
    return err / scale


def test_bp3d20_against_reference_golden(monte):
    """recon/bp3d20.cpp as shipped (65x65x360 -> 256^3, s in [125,130), sphere mask)."""
    gold = np.load(os.path.join(GOLDEN, "fdk_bp3d20.npz"))
    g = _abi.bp3d20_geom()
    f, xy, zy, st = monte.fdk(g, rand(int(gold["seed"]), (360, 65, 65)), want_zy=True)
    assert_close(f[gold["views_kept"]], gold["filtered_views"], "filtered")
    assert_close(xy[:, :, 125:130][::4, ::4, :], gold["slab_sub"], "slab")
    assert not xy[:, :, :125].any() and not xy[:, :, 130:].any()
    assert np.array_equal(zy.transpose(2, 1, 0), xy)
    assert st["launches"] >= 3 and st["voxel_updates"] == 256 * 256 * 5 * 360


def test_bp3d20_325_against_reference_golden(monte):
    gold = np.load(os.path.join(GOLDEN, "fdk_bp3d20_325.npz"))
    g = _abi.bp3d20_325_geom()
    f, xy, _, _ = monte.fdk(g, rand(int(gold["seed"]), (360, 325, 325)))
    assert_close(f[gold["views_kept"]][:, ::4, :], gold["filtered_views"], "filtered")
    assert_close(xy[:, :, 125:130][::4, ::4, :], gold["slab_sub"], "slab")


def test_bp3d20_full_slab_against_oracle(monte, oracle):
    """every voxel of the shipped slab, every filtered pixel, against the CPU restatement"""
    g = _abi.bp3d20_geom()
    proj = rand(5, (360, 65, 65))
    f_o, xy_o, _ = oracle.fdk(g, proj)
    f, xy, _, _ = monte.fdk(g, proj)
    assert_close(f, f_o, "filtered")
    assert_close(xy, xy_o, "volume")


@pytest.mark.parametrize("nu,nv,n,views,textbook", [
    (65, 65, 48, 90, False),       # reference-style square detector
    (96, 40, 40, 60, False),       # ragged: nu != nv, neither a multiple of the tile sizes
    (33, 17, 24, 45, True),        # tiny + textbook weights
    (130, 70, 56, 120, True),
])
def test_generic_geometry_against_oracle(monte, oracle, nu, nv, n, views, textbook):
    g = _abi.generic_fdk_geom(views, nu, nv, n, textbook=textbook)
    proj = rand(nu * 1000 + nv, (views, nu, nv))
    f_o, xy_o, zy_o = oracle.fdk(g, proj, want_zy=True)
    f, xy, zy, _ = monte.fdk(g, proj, want_zy=True)
    assert_close(f, f_o, "filtered")
    assert_close(xy, xy_o, "volume")
    assert np.array_equal(zy.transpose(2, 1, 0), xy)


def test_partial_roi_and_mask(monte, oracle):
    g = _abi.generic_fdk_geom(72, 65, 65, 64)
    g.s_begin, g.s_end, g.t_begin, g.t_end, g.z_begin, g.z_end = 3, 50, 7, 64, 10, 33
    g.mask_cs = g.mask_ct = g.mask_cz = 32
    g.mask_r2 = 25 * 25
    proj = rand(9, (72, 65, 65))
    _, xy_o, _ = oracle.fdk(g, proj)
    _, xy, _, _ = monte.fdk(g, proj, want_filtered=False)
    assert_close(xy, xy_o, "volume")
    assert np.array_equal(xy == 0, xy_o == 0)


def test_empty_roi_gives_zero_volume(monte):
    g = _abi.generic_fdk_geom(8, 33, 33, 16)
    g.s_begin = g.s_end = 0
    _, xy, _, _ = monte.fdk(g, rand(1, (8, 33, 33)), want_filtered=False)
    assert not xy.any()


def test_fbp2_against_reference_golden(monte):
    gold = np.load(os.path.join(GOLDEN, "fdk_fbp2.npz"))
    f, img, _ = monte.fbp2(_abi.fbp2_geom(), rand(int(gold["seed"]), (360, 65)), view_first=1)
    assert_close(f[::8], gold["filtered"], "fbp2 filtered")
    assert_close(img[::2, ::2], gold["image_sub"], "fbp2 image")


def test_linearity_of_the_whole_pipeline(monte):
    """size-independent property: FDK is linear in the projections"""
    g = _abi.generic_fdk_geom(64, 80, 48, 48)
    a, b = rand(3, (64, 80, 48)), rand(4, (64, 80, 48))
    va = monte.fdk(g, a, want_filtered=False)[1]
    vb = monte.fdk(g, b, want_filtered=False)[1]
    vab = monte.fdk(g, a + 2 * b, want_filtered=False)[1]
    assert_close(va + 2 * vb, vab, "linearity")


def test_device_api_slabs_equal_whole(monte):
    """z-slab sharding (the multi-GPU partition) reproduces the single launch bit for bit"""
    import torch
    g = _abi.generic_fdk_geom(60, 65, 65, 40)
    proj = torch.from_numpy(rand(8, (60, 65, 65))).cuda()
    filt = torch.empty(monte.fdk_filtered_shape(g), dtype=torch.float32, device="cuda")
    monte.fdk_filter_dev(g, proj, filt)
    whole = torch.empty((g.nz, g.ny, g.nx), dtype=torch.float32, device="cuda")
    monte.fdk_backproject_dev(g, filt, whole)
    parts = torch.empty_like(whole)
    for lo, hi in ((0, 7), (7, 24), (24, 40)):
        monte.fdk_backproject_dev(g, filt, parts[lo:hi], lo, hi)
    # filter by view ranges too
    filt2 = torch.zeros_like(filt)
    monte.fdk_filter_dev(g, proj, filt2, 0, 25, pad=False)
    monte.fdk_filter_dev(g, proj, filt2, 25, 60, pad=False)
    monte.fdk_pad_dev(g, filt2)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    assert torch.equal(filt, filt2)
    ref = monte.fdk(g, proj.cpu().numpy(), want_filtered=False)[1]
    assert np.array_equal(whole.cpu().numpy(), ref)


def test_host_pipeline_chunks_equal_single_launch(monte):
    """the host-buffer call streams views in 8 chunks and downloads 4 z-slabs (>= 64 views, nz >= 128);
    fp32 partial sums are reloaded exactly, so it equals the one-launch device path bit for bit"""
    import torch
    g = _abi.generic_fdk_geom(96, 48, 40, 128)
    g.s_begin, g.s_end, g.t_begin, g.t_end = 40, 88, 30, 100          # keep it small: a 48 x 70 x 128 region
    proj = rand(21, (96, 48, 40))
    f, vol, _, st = monte.fdk(g, proj)
    d_proj = torch.from_numpy(proj).cuda()
    filt = torch.empty(monte.fdk_filtered_shape(g), dtype=torch.float32, device="cuda")
    monte.fdk_filter_dev(g, d_proj, filt)
    whole = torch.empty((g.nz, g.ny, g.nx), dtype=torch.float32, device="cuda")
    monte.fdk_backproject_dev(g, filt, whole)
    torch.cuda.synchronize()
    assert np.array_equal(vol, whole.cpu().numpy())
    pitch = filt.shape[1]
    assert np.array_equal(f, filt[: 96 * 40].view(96, 40, pitch)[:, :, :48].cpu().numpy())
    assert st["launches"] > 20


@pytest.mark.parametrize("nu,nv", [(300, 37), (640, 21), (1100, 10), (2048, 3)])
def test_fft_and_direct_filter_against_oracle(monte, oracle, nu, nv, monkeypatch):
    """Wide detectors take the FFT filter (transform length 1024 / 2048 / 4096, odd nv leaves a lone
    column in the last pair); both it and the direct convolution must match the oracle's filter."""
    import torch
    g = _abi.generic_fdk_geom(3, nu, nv, 16)
    p = rand(nu + nv, (3, nu, nv)) - 0.25
    ref = oracle.fdk_filter(g, p)
    d = torch.from_numpy(p).cuda()
    got = {}
    for mode in ("direct", "fft"):
        monkeypatch.setenv("MONTE_FDK_FILTER", mode)
        f = torch.full(monte.fdk_filtered_shape(g), 7.0, dtype=torch.float32, device="cuda")
        monte.fdk_filter_dev(g, d, f)
        got[mode] = f[: 3 * nv].view(3, nv, f.shape[1])[:, :, :nu].cpu().numpy()
        assert_close(got[mode], ref, "filter (%s)" % mode)
    assert float(np.abs(got["fft"] - got["direct"]).max()) <= 1e-5 * float(np.abs(ref).max())


def test_c3_full_size(monte, oracle):
    """BASELINE config 3 at full size: 512^3 from 720 views of a 1024x768 detector.
    (a) the filter of two full-size views against the oracle; (b) the whole reconstruction on the GPU,
    probe voxels against the oracle's backprojection of the same filtered projections; (c) linearity."""
    import torch
    g = _abi.generic_fdk_geom(720, 1024, 768, 512)
    g2 = g.copy()
    g2.n_views = 2
    p2 = rand(77, (2, 1024, 768))
    d2 = torch.from_numpy(p2).cuda()
    f2 = torch.empty(monte.fdk_filtered_shape(g2), dtype=torch.float32, device="cuda")
    monte.fdk_filter_dev(g2, d2, f2)
    f2h = f2[: 2 * 768].view(2, 768, f2.shape[1])[:, :, :1024].cpu().numpy()
    assert_close(f2h, oracle.fdk_filter(g2, p2), "filter 1024x768")
    del d2, f2
    gen = torch.Generator(device="cuda").manual_seed(5)
    proj = torch.rand((720, 1024, 768), device="cuda", generator=gen)
    filt = torch.empty(monte.fdk_filtered_shape(g), dtype=torch.float32, device="cuda")
    vol = torch.empty((512, 512, 512), dtype=torch.float32, device="cuda")
    monte.fdk_filter_dev(g, proj, filt)
    monte.fdk_backproject_dev(g, filt, vol)
    dense = torch.empty((720, 768, 1024), dtype=torch.float32, device="cuda")
    monte.fdk_unpad_dev(g, filt, dense)
    torch.cuda.synchronize()
    dense_h = dense.cpu().numpy()
    vol_h = vol.cpu().numpy()
    scale = float(np.abs(vol_h).max())
    probes = [(0, 0, 0), (511, 511, 511), (256, 256, 256), (17, 400, 33), (300, 5, 470), (128, 256, 384), (500, 100, 250), (3, 255, 256)]
    for (z, t, s) in probes:
        gp = g.copy()
        gp.s_begin, gp.s_end, gp.t_begin, gp.t_end, gp.z_begin, gp.z_end = s, s + 1, t, t + 1, z, z + 1
        ref = oracle.fdk_backproject(gp, dense_h)[z, t, s]
        assert abs(float(vol_h[z, t, s]) - float(ref)) <= REL * scale, ((z, t, s), vol_h[z, t, s], ref, scale)
    # linearity at full size: FDK(2p) == 2 FDK(p) exactly in fp32 (power-of-two scaling)
    proj *= 2.0
    vol2 = torch.empty_like(vol)
    monte.fdk_filter_dev(g, proj, filt)
    monte.fdk_backproject_dev(g, filt, vol2)
    torch.cuda.synchronize()
    assert torch.equal(vol2, vol * 2.0)


def test_bad_arguments_are_reported_not_fatal(monte):
    g = _abi.generic_fdk_geom(4, 16, 16, 8)
    g.s_end = 99
    with pytest.raises(monte.MonteError, match="ROI"):
        monte.fdk(g, rand(0, (4, 16, 16)))
    # detectors wider than the FFT filter's longest transform fall back to the direct convolution;
    # forcing the FFT there is an error, not a crash
    import torch
    gw = _abi.generic_fdk_geom(1, 2100, 2, 8)
    d = torch.rand((1, 2100, 2), device="cuda")
    f = torch.zeros(monte.fdk_filtered_shape(gw), dtype=torch.float32, device="cuda")
    monte.fdk_filter_dev(gw, d, f)
    torch.cuda.synchronize()
    assert float(f[:2, :2100].abs().max()) > 0
    os.environ["MONTE_FDK_FILTER"] = "fft"
    try:
        with pytest.raises(monte.MonteError, match="too wide"):
            monte.fdk_filter_dev(gw, d, f)
    finally:
        del os.environ["MONTE_FDK_FILTER"]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["10", "11"])
def test_tma_staged_backprojector_gives_the_bits_of_the_default_kernel(monte, variant):
    """MONTE_BP_VARIANT=10/11: the footprint of a 16x16x16 brick staged in shared memory by TMA (cp.async.bulk.tensor.2d,
    3 stages, full/empty mbarriers) instead of gathered through L1 -- same arithmetic on the same texels, so the volume is
    bit-identical, including ROI / mask handling, ragged sizes and the view-chunk continuation"""
    import torch
    for n_views, nu, nv, n, roi in ((31, 96, 40, 40, False), (48, 65, 65, 32, True), (90, 300, 200, 96, False)):
        g = _abi.generic_fdk_geom(n_views, nu, nv, n)
        if roi:
            g.s_begin, g.s_end, g.t_begin, g.t_end, g.z_begin, g.z_end = 3, n - 5, 2, n - 1, 4, n - 6
            g.mask_cs = g.mask_ct = g.mask_cz = n // 2
            g.mask_r2 = (n // 2 - 2) ** 2
        gen = torch.Generator(device="cuda").manual_seed(n_views)
        proj = torch.rand((n_views, nu, nv), device="cuda", generator=gen)
        filt = torch.zeros(monte.fdk_filtered_shape(g), device="cuda")
        monte.fdk_filter_dev(g, proj, filt)
        vols = {}
        for var in ("0", variant):
            os.environ["MONTE_BP_VARIANT"] = var
            try:
                v = torch.full((n, n, n), float("nan"), device="cuda")
                monte.fdk_backproject_dev(g, filt, v)
                # ... and in two view pieces that continue the stored partial sums
                w = torch.full((n, n, n), float("nan"), device="cuda")
                monte.fdk_backproject_views_dev(g, filt, w, 0, n, 0, n_views // 3, False)
                monte.fdk_backproject_views_dev(g, filt, w, 0, n, n_views // 3, n_views, True)
                torch.cuda.synchronize()
            finally:
                del os.environ["MONTE_BP_VARIANT"]
            assert torch.equal(v, w)
            vols[var] = v
        assert bool(torch.isfinite(vols["0"]).all()) and torch.equal(vols["0"], vols[variant])


def test_thin_slab_z_block_streams_are_bit_identical(tmp_path):
    """a thin z-slab (2..8 z-blocks) backprojected in several view chunks runs every z-block on its own stream so that
    the tail of one launch is filled by the next z-block's CTAs (multi-GPU slabs): same bits as the single-stream
    launches and as the whole-volume call.  Own process: the view-chunk size is read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
from monte_b200 import _abi, api
api.init(0)
g = _abi.generic_fdk_geom(26, 72, 56, 64)
proj = torch.rand((26, 72, 56), device="cuda")
filt = torch.zeros(api.fdk_filtered_shape(g), device="cuda")
api.fdk_filter_dev(g, proj, filt)
whole = torch.zeros((64, 64, 64), device="cuda")
api.fdk_backproject_dev(g, filt, whole)
outs = []
for zs in ("1", "0"):
    os.environ["MONTE_BP_ZSTREAMS"] = zs
    slab = torch.full((43, 64, 64), float("nan"), device="cuda")
    api.fdk_backproject_dev(g, filt, slab, 9, 52)          # 4 z-blocks (one ragged at each end), 26 views in chunks of 4
    torch.cuda.synchronize()
    outs.append(slab)
assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], whole[9:52]) and bool(torch.isfinite(outs[0]).all())
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MONTE_BP_VCHUNK="4")
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("textbook", [False, True])
def test_two_d_fan_beam_is_the_nv1_case_of_the_same_kernels(monte, oracle, textbook):
    """a 2-D fan-beam reconstruction through the cone-beam filter and backprojector (nv = 1, one slice in
    the central plane): against the oracle on random data, and -- Feldkamp weights -- a disc's attenuation coefficient
    recovered from its analytic fan-beam line integrals"""
    g = _abi.fan_beam_2d_geom(90, 96, 48, textbook=textbook)
    proj = rand(5, (90, 96, 1))
    f, vol, _, st = monte.fdk(g, proj)
    fo, vo, _ = oracle.fdk(g, proj)
    assert vol.shape == (1, 48, 48) and np.abs(vo).max() > 0
    assert np.abs(f - fo).max() <= REL * np.abs(fo).max() and np.abs(vol - vo).max() <= REL * np.abs(vo).max()
    if not textbook:
        return
    g = _abi.fan_beam_2d_geom(360, 192, 64)
    r, mu = 8.0, 0.2
    u = g.half_u - g.du * (np.arange(g.nu) + 0.0)              # detector coordinate of column iu (bp3d20.cpp:40: 16.25 - pitch * zeta)
    t = g.dso * u / np.sqrt(u * u + g.dsd * g.dsd)             # distance of the ray from the rotation centre
    chord = 2.0 * mu * np.sqrt(np.maximum(r * r - t * t, 0.0))
    proj = np.broadcast_to(chord.astype(np.float32)[None, :, None], (360, g.nu, 1)).copy()
    _, vol, _, _ = monte.fdk(g, proj, want_filtered=False)
    c = vol[0, 24:40, 24:40]
    assert abs(c.mean() - mu) < 0.02 * mu, c.mean()
    assert abs(vol[0, 32, 7]) < 0.05 * mu                       # 10 cm from the centre: outside the disc, inside the field of view


@pytest.mark.parametrize("chunks,tail", [(16, 3), (8, 8), (5, 1), (1, 1)])
def test_host_pipeline_chunking_is_bit_identical(monte, chunks, tail):
    """monte_gpu_fdk uploads, filters and backprojects the views in chunks and finishes the last `tail` chunks slab by slab
    while finished slabs go home: any chunking gives the bits of the device-resident stages"""
    import os
    g = _abi.generic_fdk_geom(41, 40, 36, 128)
    g.s_begin, g.s_end, g.t_begin, g.t_end = 58, 70, 60, 66          # a thin column of the 128^3 volume (four z-slabs)
    proj = rand(9, (41, 40, 36))
    keep = {k: os.environ.get(k) for k in ("MONTE_FDK_CHUNKS", "MONTE_FDK_TAIL")}
    try:
        os.environ["MONTE_FDK_CHUNKS"], os.environ["MONTE_FDK_TAIL"] = "1", "1"
        f0, v0, z0, _ = monte.fdk(g, proj, want_zy=True)
        os.environ["MONTE_FDK_CHUNKS"], os.environ["MONTE_FDK_TAIL"] = str(chunks), str(tail)
        f1, v1, z1, st = monte.fdk(g, proj, want_zy=True)
    finally:
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert np.array_equal(f0, f1) and np.array_equal(v0, v1) and np.array_equal(z0, z1)
    assert np.abs(v0).max() > 0 and not v0[:, :, :58].any()


def test_backproject_from_segment_buffers_equals_one_buffer(monte):
    """monte_gpu_fdk_backproject_peers_dev (the one-process-per-GPU exchange: every rank's filtered views stay in its own
    buffer, the backprojector gathers the band out of them): three "ranks" as three buffers of one device, everything a
    rank did not filter is NaN -- same bits as the padded single buffer"""
    import torch
    g = _abi.generic_fdk_geom(37, 56, 40, 48)
    proj = torch.from_numpy(rand(11, (37, 56, 40))).cuda()
    filt = torch.zeros(monte.fdk_filtered_shape(g), device="cuda")
    monte.fdk_filter_dev(g, proj, filt)
    cuts = [0, 9, 9, 30, 37]                                        # four segments, one of them empty
    bufs = []
    for a, b in zip(cuts, cuts[1:]):
        f = torch.full(monte.fdk_filtered_shape(g), float("nan"), device="cuda")
        if b > a:
            monte.fdk_filter_dev(g, proj, f, a, b, pad=False)
        bufs.append(f)
    for z_lo, z_hi in ((0, 48), (5, 29), (32, 48)):
        want = torch.empty((z_hi - z_lo, 48, 48), device="cuda")
        monte.fdk_backproject_dev(g, filt, want, z_lo, z_hi)
        got = torch.full_like(want, float("nan"))
        monte.fdk_backproject_peers_dev(g, [f.data_ptr() for f in bufs], cuts[1:], got, z_lo, z_hi)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (z_lo, z_hi)
    with pytest.raises(Exception, match="cover"):
        monte.fdk_backproject_peers_dev(g, [bufs[0].data_ptr()], [9], got, 32, 48)
    # the IPC plumbing inside one process: export gives the allocation's handle and the offset of the pointer in it
    h, off = monte.ipc_export(filt[3:])
    h0, off0 = monte.ipc_export(filt)
    assert len(h) == 64 and h == h0 and off - off0 == 3 * filt.shape[1] * 4
