"""CPU tests of the FDK kernels and their host code under SIMT emulation (tests/emu).

libmonte_gpu's own .cu sources are compiled by g++ against tests/emu/cuda_runtime.h (CTAs run one after the
other, threads are fibers that switch at barriers) and driven through the same C ABI as on the GPU.  The
bodies of the GPU parity tests are reused where they only need host buffers, so what `-m gpu` checks on a
B200 is checked here for indexing, synchronisation and host-side chunking logic -- not for speed, and fp32
results can differ from the GPU's in the last ulp (FMA contraction, MUFU approximations).
This is test infrastructure: the product has no CPU path (tests/test_abi.py::test_no_cpu_fallback_without_a_device).
"""
import os

import numpy as np
import pytest

import test_fdk_gpu as G
from monte_b200 import _abi

rand, assert_close = G.rand, G.assert_close


def _Dev(a):
    """numpy array -> the "device buffer" handle the *_dev wrappers of api.py expect (tests/emu/build.py)"""
    return _Dev.cls(a)


@pytest.fixture(autouse=True)
def _bind_dev(monte_emu):
    _Dev.cls = monte_emu.Dev


@pytest.mark.parametrize("nu,nv,n,views,textbook", [
    (65, 65, 32, 48, False),
    (96, 40, 40, 30, False),       # ragged: nu != nv, neither a multiple of the tile sizes
    (33, 17, 24, 45, True),        # tiny + textbook weights
])
def test_emu_generic_geometry_against_oracle(monte_emu, oracle, nu, nv, n, views, textbook):
    G.test_generic_geometry_against_oracle(monte_emu, oracle, nu, nv, n, views, textbook)


def test_emu_partial_roi_and_mask(monte_emu, oracle):
    G.test_partial_roi_and_mask(monte_emu, oracle)


def test_emu_empty_roi_gives_zero_volume(monte_emu):
    G.test_empty_roi_gives_zero_volume(monte_emu)


def test_emu_fbp2_against_reference_golden(monte_emu):
    G.test_fbp2_against_reference_golden(monte_emu)


def test_emu_bp3d20_shipped_slab_subset_against_golden(monte_emu):
    """recon/bp3d20.cpp's geometry (65x65x360 -> 256^3, sphere mask), a 2-column x 64-slice piece of the shipped
    slab s in [125,130): against the volume the UNMODIFIED reference binary wrote (tests/golden)."""
    gold = np.load(os.path.join(G.GOLDEN, "fdk_bp3d20.npz"))
    g = _abi.bp3d20_geom()
    g.s_begin, g.s_end, g.z_begin, g.z_end = 125, 127, 96, 160
    f, xy, _, _ = monte_emu.fdk(g, rand(int(gold["seed"]), (360, 65, 65)))
    assert_close(f[gold["views_kept"]], gold["filtered_views"], "filtered")
    sub = xy[:, :, 125:130][::4, ::4, :]                      # golden: every 4th z and t of the slab
    zs = np.arange(0, 256, 4)
    inside = (zs >= 96) & (zs < 160)
    scale = float(np.abs(gold["slab_sub"]).max())
    assert np.abs(sub[inside][:, :, :2] - gold["slab_sub"][inside][:, :, :2]).max() <= G.REL * scale
    assert not xy[:96].any() and not xy[160:].any() and not xy[:, :, 127:].any()


def test_emu_linearity_of_the_whole_pipeline(monte_emu):
    G.test_linearity_of_the_whole_pipeline(monte_emu)


def test_emu_device_api_slabs_equal_whole(monte_emu):
    """z-slab sharding (the multi-GPU partition) and view-range filtering reproduce the single launch bit for bit"""
    m = monte_emu
    g = _abi.generic_fdk_geom(40, 65, 65, 40)
    proj = rand(8, (40, 65, 65))
    filt = np.full(m.fdk_filtered_shape(g), np.nan, np.float32)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt))
    whole = np.full((g.nz, g.ny, g.nx), np.nan, np.float32)
    m.fdk_backproject_dev(g, _Dev(filt), _Dev(whole))
    parts = np.full_like(whole, np.nan)
    for lo, hi in ((0, 7), (7, 24), (24, 40)):
        m.fdk_backproject_dev(g, _Dev(filt), _Dev(parts[lo:hi]), lo, hi)
    filt2 = np.zeros_like(filt)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt2), 0, 25, pad=False)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt2), 25, 40, pad=False)
    m.fdk_pad_dev(g, _Dev(filt2))
    assert np.array_equal(whole, parts)
    assert np.array_equal(filt, filt2)
    assert np.array_equal(whole, m.fdk(g, proj, want_filtered=False)[1])
    # views fed in ascending pieces continue the fp32 partial sums exactly (the pipelined multi-GPU exchange)
    piece = np.full_like(whole, np.nan)
    for i, (lo, hi) in enumerate(((0, 13), (13, 14), (14, 40))):
        m.fdk_backproject_views_dev(g, _Dev(filt), _Dev(piece), 0, g.nz, lo, hi, i > 0)
    assert np.array_equal(whole, piece)


def test_emu_backproject_from_segment_buffers_equals_one_buffer(monte_emu):
    """monte_gpu_fdk_backproject_peers_dev: the views stay in the buffers of the "ranks" that filtered them (everything else
    NaN), the backprojector's pair conversion gathers the band out of them -- the bits of the padded single buffer"""
    m = monte_emu
    g = _abi.generic_fdk_geom(37, 56, 40, 48)
    proj = rand(11, (37, 56, 40))
    filt = np.zeros(m.fdk_filtered_shape(g), np.float32)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt))
    cuts = [0, 9, 9, 30, 37]
    bufs = []
    for a, b in zip(cuts, cuts[1:]):
        f = np.full(m.fdk_filtered_shape(g), np.nan, np.float32)
        if b > a:
            m.fdk_filter_dev(g, _Dev(proj), _Dev(f), a, b, pad=False)
        bufs.append(f)
    for z_lo, z_hi in ((0, 48), (5, 29), (32, 48)):
        want = np.zeros((z_hi - z_lo, 48, 48), np.float32)
        m.fdk_backproject_dev(g, _Dev(filt), _Dev(want), z_lo, z_hi)
        got = np.full_like(want, np.nan)
        m.fdk_backproject_peers_dev(g, [f.ctypes.data for f in bufs], cuts[1:], _Dev(got), z_lo, z_hi)
        assert np.array_equal(got, want), (z_lo, z_hi)
    with pytest.raises(m.MonteError, match="cover"):
        m.fdk_backproject_peers_dev(g, [bufs[0].ctypes.data], [9], _Dev(got), 32, 48)
    h, off = m.ipc_export(_Dev(filt))
    p = m.ipc_open(h, off)
    assert p == filt.ctypes.data
    m.ipc_close(p)
    with pytest.raises(m.MonteError, match="ipc_close"):
        m.ipc_close(p)


def test_emu_band_limited_rows_are_all_a_slab_reads(monte_emu):
    """monte_gpu_fdk_slab_rows: with every row outside the reported band (and rows 0..3) set to NaN the slab
    still equals the full-data volume -- the contract the band-limited multi-GPU exchange relies on"""
    m = monte_emu
    g = _abi.generic_fdk_geom(36, 48, 96, 48)
    proj = rand(3, (36, 48, 96))
    filt = np.zeros(m.fdk_filtered_shape(g), np.float32)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt))
    whole = np.zeros((g.nz, g.ny, g.nx), np.float32)
    m.fdk_backproject_dev(g, _Dev(filt), _Dev(whole))
    for lo, hi in ((0, 16), (16, 32), (32, 48)):
        r0, r1 = m.fdk_slab_rows(g, lo, hi)
        assert 0 <= r0 <= r1 <= g.nv and r1 - r0 < g.nv
        holed = filt.copy()
        rows = holed[: g.n_views * g.nv].reshape(g.n_views, g.nv, -1)
        keep = np.zeros(g.nv, bool)
        keep[r0:r1] = True
        keep[:4] = True
        rows[:, ~keep, :] = np.nan
        slab = np.zeros((hi - lo, g.ny, g.nx), np.float32)
        m.fdk_backproject_dev(g, _Dev(holed), _Dev(slab), lo, hi)
        assert np.array_equal(slab, whole[lo:hi])


def test_emu_host_pipeline_chunks_equal_single_launch(monte_emu):
    """>= 64 views and nz >= 128 switch the host-buffer call to 8 view chunks and 4 z-slabs; the result equals
    the one-launch device path bit for bit"""
    m = monte_emu
    g = _abi.generic_fdk_geom(64, 48, 40, 128)
    g.s_begin, g.s_end, g.t_begin, g.t_end = 56, 72, 50, 66            # a 16 x 16 x 128 region
    proj = rand(21, (64, 48, 40))
    f, vol, _, st = m.fdk(g, proj)
    filt = np.full(m.fdk_filtered_shape(g), np.nan, np.float32)
    m.fdk_filter_dev(g, _Dev(proj), _Dev(filt))
    whole = np.full((g.nz, g.ny, g.nx), np.nan, np.float32)
    m.fdk_backproject_dev(g, _Dev(filt), _Dev(whole))
    assert np.array_equal(vol, whole)
    assert np.array_equal(f, filt[: 64 * 40].reshape(64, 40, -1)[:, :, :48])
    assert st["launches"] > 20


@pytest.mark.parametrize("nu,nv", [(300, 11), (640, 5), (1100, 3)])
def test_emu_fft_and_direct_filter_against_oracle(monte_emu, oracle, nu, nv, monkeypatch):
    """the FFT filter (transform length 1024 / 2048 / 4096; odd nv leaves a lone column in the last pair) and
    the direct convolution against the oracle's filter"""
    m = monte_emu
    g = _abi.generic_fdk_geom(2, nu, nv, 16)
    p = rand(nu + nv, (2, nu, nv)) - 0.25
    ref = oracle.fdk_filter(g, p)
    got = {}
    for mode in ("direct", "fft"):
        monkeypatch.setenv("MONTE_FDK_FILTER", mode)
        f = np.full(m.fdk_filtered_shape(g), 7.0, np.float32)
        m.fdk_filter_dev(g, _Dev(p), _Dev(f))
        got[mode] = f[: 2 * nv].reshape(2, nv, -1)[:, :, :nu]
        assert_close(got[mode], ref, "filter (%s)" % mode)
    assert float(np.abs(got["fft"] - got["direct"]).max()) <= 1e-5 * float(np.abs(ref).max())


def test_emu_bad_arguments_are_reported_not_fatal(monte_emu, monkeypatch):
    m = monte_emu
    g = _abi.generic_fdk_geom(4, 16, 16, 8)
    g.s_end = 99
    with pytest.raises(m.MonteError, match="ROI"):
        m.fdk(g, rand(0, (4, 16, 16)))
    gw = _abi.generic_fdk_geom(1, 2100, 2, 8)            # wider than the longest transform: direct convolution
    d, f = rand(1, (1, 2100, 2)), np.zeros(m.fdk_filtered_shape(gw), np.float32)
    m.fdk_filter_dev(gw, _Dev(d), _Dev(f))
    assert float(np.abs(f[:2, :2100]).max()) > 0
    monkeypatch.setenv("MONTE_FDK_FILTER", "fft")
    with pytest.raises(m.MonteError, match="too wide"):
        m.fdk_filter_dev(gw, _Dev(d), _Dev(f))


def test_emu_shutdown_and_reinit_rebuild_every_cache(monte_emu, oracle):
    """monte_gpu_shutdown frees the cached device buffers, events and per-context attributes; a new
    monte_gpu_init starts from scratch (host-buffer FDK with the 8-chunk pipeline, then MC) and gives the same bits"""
    from monte_b200 import scenes
    m = monte_emu
    g = _abi.generic_fdk_geom(64, 40, 24, 128)
    g.s_begin, g.s_end, g.t_begin, g.t_end = 60, 68, 60, 68
    proj = rand(2, (64, 40, 24))
    lab = scenes.cylinder_phantom(17, 2.0)
    mg = scenes.mc_geom(9, 32.5 / 9, n_views=1)
    args = (mg, scenes.volume_for(lab, 2.0), lab, scenes.make_xs(), scenes.mono_spectrum(), 20, 3)
    f1, v1, _, _ = m.fdk(g, proj)
    a0, a5, _ = m.simulate(*args)
    m.shutdown()
    with pytest.raises(m.MonteError, match="monte_gpu_init"):
        m.fdk(g, proj)
    m.init(0)
    f2, v2, _, _ = m.fdk(g, proj)
    b0, b5, _ = m.simulate(*args)
    assert np.array_equal(f1, f2) and np.array_equal(v1, v2) and np.array_equal(a0, b0) and np.array_equal(a5, b5)


@pytest.mark.parametrize("textbook", [False, True])
def test_emu_two_d_fan_beam_is_the_nv1_case_of_the_same_kernels(monte_emu, oracle, textbook):
    G.test_two_d_fan_beam_is_the_nv1_case_of_the_same_kernels(monte_emu, oracle, textbook)


@pytest.mark.parametrize("chunks,tail", [(16, 3), (8, 8), (5, 1)])
def test_emu_host_pipeline_chunking_is_bit_identical(monte_emu, chunks, tail):
    G.test_host_pipeline_chunking_is_bit_identical(monte_emu, chunks, tail)


def test_emu_smoke_path(monte_emu):
    """__graft_entry__.smoke()'s own checks (FDK, FFT filter, MC coupled with the oracle, projector, the optional
    transport modes), run on the emulated library; leaves the binding initialised for the tests that follow"""
    import os
    import __graft_entry__ as ge
    ge._smoke(monte_emu)
    monte_emu.init(0)
    os.environ["MONTE_EMU_DEVICES"] = "2"                # ... and once more with two "devices": the 2-device check of smoke()
    try:
        ge._smoke(monte_emu)
    finally:
        del os.environ["MONTE_EMU_DEVICES"]
        monte_emu.init(0)


@pytest.mark.slow
def test_emu_address_sanitizer_clean():
    """tests/emu/asan_check.py: kernels and host code on the ASan build of the emulated library -- no out-of-bounds
    access of any "device" buffer (the CPU counterpart of compute-sanitizer memcheck)"""
    import subprocess
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "emu", "asan_check.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert "ERROR: AddressSanitizer" not in out, out[-4000:]
    assert r.returncode == 0 and "ASAN RUN COMPLETE" in out, out[-4000:]


def test_emu_cbct_fdk_driver_fbp2_end_to_end(monte_emu, tmp_path):
    """the recon main() replacement (monte_b200/host/cbct_fdk.cpp) linked against the emulated library, in its fbp2
    role: reads the reference's input file, writes its output files; against the golden of the unmodified
    recon/fbp2.cpp binary"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emu_dir = os.path.join(root, "tests", "emu", "_build")
    exe = os.path.join(str(tmp_path), "cbct_fdk_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "monte_b200", "host", "cbct_fdk.cpp"), "-o", exe,
                           "-L" + emu_dir, "-lmonte_gpu_emu", "-Wl,-rpath," + emu_dir])
    gold = np.load(os.path.join(G.GOLDEN, "fdk_fbp2.npz"))
    d = str(tmp_path)
    rand(int(gold["seed"]), (360, 65)).tofile(os.path.join(d, "map5_20_2e5.raw"))          # fbp2.cpp:27
    out = subprocess.run([exe, "fbp2", "map5_20_2e5.raw", "t"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         check=True).stdout.decode()
    assert "filtered" in out
    img = np.fromfile(os.path.join(d, "xy_t.raw"), np.float32).reshape(256, 256)
    f = np.fromfile(os.path.join(d, "map_t.raw"), np.float32).reshape(360, 65)
    assert_close(f[::8], gold["filtered"], "fbp2 filtered")
    assert_close(img[::2, ::2], gold["image_sub"], "fbp2 image")
    # a missing input file is reported, not fatal in any other way
    r = subprocess.run([exe, "fbp2", "nope.raw"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "failed to read" in r.stderr
