"""Seeded random sweeps of the C ABI against the oracle on the emulated library (tests/emu): ragged and degenerate
FDK geometries (non-cubic volumes, partial ROIs, masks, both weight and coordinate modes, off-centre volumes, short
scans) and MC scenes (odd detector sizes, view / photon sub-ranges, every source / detector / coherent / tracking
mode, 1..3 materials).  The same parity bars as the GPU tests: FDK 1e-4 of max|reference| on every element, MC
fates history by history.  Test infrastructure only: this is the CPU way to cover many edge cases per second of
GPU time saved; the fixed-size GPU tests remain the parity proof on hardware.
"""
import math

import numpy as np
import pytest

from monte_b200 import _abi, scenes

REL = 1e-4


def _fdk_case(rng):
    nu, nv = int(rng.integers(9, 90)), int(rng.integers(3, 60))
    views = int(rng.integers(3, 40))
    g = _abi.generic_fdk_geom(views, nu, nv, 8, textbook=bool(rng.integers(0, 2)))
    g.nx, g.ny, g.nz = int(rng.integers(3, 40)), int(rng.integers(3, 40)), int(rng.integers(1, 40))
    g.vox = float(rng.uniform(0.3, 1.2))
    g.x0 = -0.5 * g.nx * g.vox + float(rng.uniform(-2, 2))          # off-centre volumes too
    g.y0 = 0.5 * g.ny * g.vox + float(rng.uniform(-2, 2))
    g.z0 = 0.5 * g.nz * g.vox + float(rng.uniform(-3, 3))
    g.full_roi()
    if rng.integers(0, 2):
        a, b = sorted(rng.integers(0, g.nx + 1, 2)); g.s_begin, g.s_end = int(a), int(b)
        a, b = sorted(rng.integers(0, g.ny + 1, 2)); g.t_begin, g.t_end = int(a), int(b)
        a, b = sorted(rng.integers(0, g.nz + 1, 2)); g.z_begin, g.z_end = int(a), int(b)
    if rng.integers(0, 3) == 0:
        g.mask_cs, g.mask_ct, g.mask_cz = g.nx // 2, g.ny // 2, g.nz // 2
        g.mask_r2 = int(rng.integers(1, 400))
    g.coord_mode = int(rng.integers(0, 2))
    g.angle0_deg = float(rng.uniform(-30, 30))
    if rng.integers(0, 3) == 0:
        g.angle_step_deg = float(rng.uniform(0.5, 3.0))               # short scan
    return g


@pytest.mark.parametrize("seed", range(48))
def test_emu_fuzz_fdk_against_oracle(monte_emu, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    g = _fdk_case(rng)
    proj = rng.random((g.n_views, g.nu, g.nv), dtype=np.float32) - 0.3
    want_zy = bool(rng.integers(0, 2))
    f_o, xy_o, zy_o = oracle.fdk(g, proj, want_zy=want_zy)
    f, xy, zy, _ = monte_emu.fdk(g, proj, want_zy=want_zy)
    for got, ref, what in ((f, f_o, "filtered"), (xy, xy_o, "volume")):
        scale = float(np.abs(ref).max())
        err = float(np.abs(got.astype(np.float64) - ref).max())
        assert err <= REL * scale + 1e-30, (what, err, scale, [getattr(g, k) for k, _ in g._fields_])
    assert np.array_equal(xy == 0, xy_o == 0) or np.abs(xy[(xy == 0) != (xy_o == 0)]).max() <= REL * np.abs(xy_o).max()
    if want_zy:
        assert np.array_equal(zy.transpose(2, 1, 0), xy)


def _mc_case(rng):
    n = int(rng.choice([17, 25, 33]))
    pitch = 33.0 / n * float(rng.uniform(0.6, 1.0))
    mats = [("h2o",), ("h2o", "ca"), ("h2o", "ca", "pmma")][int(rng.integers(0, 3))]
    lab = scenes.cylinder_phantom(n, pitch, radius=0.3 * n * pitch, half_len=0.35 * n * pitch,
                                  rods=len(mats) > 1, rod_r=0.05 * n * pitch, rod_ring=0.15 * n * pitch)
    if len(mats) == 3:
        lab[lab == 1] = np.where(rng.random((lab == 1).sum()) < 0.5, 1, 3).astype(np.uint8)   # salt-and-pepper third material
    # square detectors: monte_mc_geom has one pixel size and one half-extent like the reference (detector_height
    # 16.25 on both axes, CBCT_real325im.cu:459), and its bin formula (:574-575) is centred, n * pixel = 2 * half
    ny = nx = int(rng.integers(3, 14))
    g = scenes.mc_geom(0, 32.5 / nx, n_views=int(rng.integers(1, 5)), ny=ny, nx=nx,
                       source_mode=int(rng.integers(0, 2)), max_scatter=int(rng.integers(0, 7)))
    g.angle0_deg, g.angle_step_deg = float(rng.uniform(0, 360)), float(rng.uniform(5, 120))
    g.detector_mode = int(rng.integers(0, 2))
    xs = scenes.make_xs(mats)
    if rng.integers(0, 2):
        g.coherent_mode = _abi.COHERENT_FORMFACTOR
        scenes.add_formfactors(xs, mats)
    vol = scenes.volume_for(lab, pitch, tight=bool(rng.integers(0, 2)))
    vol.tracking_mode = int(rng.integers(0, 5))
    vol.clearance_cell_log2 = int(rng.integers(0, 4))
    if rng.integers(0, 2):
        spec, keep = scenes.kramers_spectrum(kvp=float(rng.choice([80.0, 120.0])))
    else:
        spec, keep = scenes.mono_spectrum(float(rng.uniform(25, 150))), None
    # majorant over the materials present only (monte_mc_volume.majorant_mode); half of those cases lose their calcium,
    # so the majorant really drops and the clearance modes meet a heavy material that does not occur
    vol.majorant_mode = int(rng.integers(0, 2))
    if vol.majorant_mode and len(mats) > 1 and rng.integers(0, 2):
        lab[lab == 2] = 1
    # a quarter of the cases on the ring detector (source at the origin, ny angular x nx axial bins on a
    # cylinder around the clip box), with whatever tracking and coherent mode the case drew
    if rng.integers(0, 4) == 0:
        g.detector_shape = _abi.DETECTOR_RING
        corner = math.hypot(max(abs(vol.clip_lo[0]), abs(vol.clip_hi[0])), max(abs(vol.clip_lo[1]), abs(vol.clip_hi[1])))
        g.ring_radius = corner * float(rng.uniform(1.05, 2.0))
    return g, vol, lab, xs, spec, keep


@pytest.mark.parametrize("seed", range(48))
def test_emu_fuzz_mc_fates_against_oracle(monte_emu, oracle, seed):
    m = monte_emu
    rng = np.random.default_rng(2000 + seed)
    g, vol, lab, xs, spec, keep = _mc_case(rng)
    per, sd = int(rng.integers(1, 40)), int(rng.integers(0, 2 ** 40))
    view = int(rng.integers(0, g.n_views))
    # what the kernel will do for AUTO, told to the oracle explicitly
    mode, cl = vol.tracking_mode, vol.clearance_cell_log2
    if mode == _abi.TRACK_AUTO:
        mode, cl, _ = m.resolve_tracking(xs, spec)
    ovol = _abi.McVolume.from_buffer_copy(vol)
    ovol.tracking_mode, ovol.clearance_cell_log2 = mode, cl
    opts = oracle.mc_opts(oracle.RNG_PHILOX, seed=sd)
    if mode in (_abi.TRACK_CLEARANCE, _abi.TRACK_ADAPTIVE, _abi.TRACK_DIRECTIONAL) and xs.n_materials > 1:
        grid, heavy = m.clearance_grid(ovol, lab, xs)
        opts, keepg = oracle.with_clearance(opts, grid, heavy)
    sc = m.Scene(g, vol, lab, xs, spec)
    f_k, e_k = sc.fates(view, per, sd)
    sc.close()
    i0, i5, res, f_o, e_o = oracle.mc_run(g, ovol, lab, oracle.tables_from_xs(xs), spec, opts, per,
                                          views=(view, view + 1), want_fates=True)
    same = f_k == f_o
    assert same.mean() >= 0.995, (same.mean(), same.size)
    assert np.allclose(e_k[same], e_o[same], rtol=2e-5)
    # tallies of the same view through the host-buffer call, photons split in two ranges
    cut = int(rng.integers(0, per + 1))
    a0, a5, sa = m.simulate(g, vol, lab, xs, spec, per, sd, views=(view, view + 1), n_range=(0, cut))
    b0, b5, sb = m.simulate(g, vol, lab, xs, spec, per, sd, views=(view, view + 1), n_range=(cut, per))
    n = g.ny * g.nx * per
    assert sa["histories"] + sb["histories"] == n == res["histories"]
    scale = 16 * 200 if g.detector_mode else 1
    assert np.abs((a0 + b0)[view].astype(np.int64) - i0[view]).sum() <= 0.005 * n * scale + 2 * scale * (~same).sum()
    assert np.abs((a5 + b5)[view].astype(np.int64) - i5[view]).sum() <= 0.005 * n * scale + 2 * scale * (~same).sum()
    other = [v for v in range(g.n_views) if v != view]
    assert not (a0 + b0)[other].any() and not (a5 + b5)[other].any()


@pytest.mark.parametrize("seed", range(16))
def test_emu_fuzz_projector_against_oracle(monte_emu, oracle, seed):
    """deterministic primary projection (1e-4 of the maximum, north_star) for random phantoms, detector sizes,
    view ranges, angles (incl. axis-aligned ones, where whole rays run along voxel faces of even-sized volumes only)
    and energies"""
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.choice([9, 17, 25, 33]))
    pitch = float(rng.uniform(0.4, 1.5))
    mats = [("h2o",), ("h2o", "ca"), ("h2o", "ca", "pmma")][int(rng.integers(0, 3))]
    if seed % 2:                                                                    # voxel noise incl. a label above n_materials
        lab = rng.integers(0, len(mats) + 2, size=(n, n, n)).astype(np.uint8)
        lab[rng.random(lab.shape) < 0.5] = 0
    else:                                                                           # large homogeneous regions
        lab = scenes.cylinder_phantom(n, pitch, radius=0.4 * n * pitch, half_len=0.3 * n * pitch, rods=len(mats) > 1,
                                      rod_r=0.08 * n * pitch, rod_ring=0.2 * n * pitch)
        lab[: n // 3, : n // 2, :] = np.where(lab[: n // 3, : n // 2, :] > 0, len(mats), 0)   # a block of the last material
    nd = int(rng.integers(2, 40))
    g = scenes.mc_geom(nd, 32.5 / nd, n_views=int(rng.integers(1, 7)))
    g.angle0_deg = float(rng.choice([0.0, 90.0, 45.0, rng.uniform(0, 360)]))
    g.angle_step_deg = float(rng.choice([90.0, rng.uniform(1, 100)]))
    vol = scenes.volume_for(lab, pitch, tight=bool(rng.integers(0, 2)))
    xs = scenes.make_xs(mats)
    keV = float(rng.uniform(15, 190))
    a, b = sorted(rng.integers(0, g.n_views + 1, 2))
    views = (int(a), int(b)) if b > a else None
    got = monte_emu.project_primary(g, vol, lab, xs, keV, views=views)
    ref = oracle.project_primary(g, vol, lab, oracle.tables_from_xs(xs), keV, views=views)
    scale = float(np.abs(ref).max())
    assert float(np.abs(got.astype(np.float64) - ref).max()) <= REL * scale + 1e-12, (scale, n, pitch, nd)


@pytest.mark.parametrize("macro", ["1", "2", "3"])
def test_emu_fuzz_projector_macro_cells(macro, oracle):
    """MONTE_PROJ_MACRO=n (homogeneous macro-cells of 2^n voxels crossed in one segment; latched per process, hence
    the subprocess): the same sweep and the fixed projector test must hold at the same 1e-4 bar"""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(here, "test_emu_fuzz.py"), os.path.join(here, "test_emu_mc.py"),
                        "-k", "fuzz_projector_against_oracle or emu_project_primary"],
                       env=dict(os.environ, MONTE_PROJ_MACRO=macro), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])


@pytest.mark.parametrize("seed", range(12))
def test_emu_fuzz_fbp2_against_oracle(monte_emu, oracle, seed):
    """2-D fan-beam FBP (recon/fbp2.cpp: truncated detector index, views from view_first) for random sinogram widths,
    image sizes, ROIs and first views"""
    rng = np.random.default_rng(4000 + seed)
    g = _abi.fbp2_geom()
    g.nu = int(rng.integers(5, 120))
    g.n_views = int(rng.integers(3, 90))
    g.du = 32.5 / g.nu
    g.half_u = 0.5 * g.nu * g.du
    g.angle_step_deg = float(rng.choice([1.0, 360.0 / g.n_views]))
    g.nx, g.ny = int(rng.integers(4, 70)), int(rng.integers(4, 70))
    g.vox = float(rng.uniform(0.15, 0.6))
    g.x0, g.y0 = -0.5 * g.nx * g.vox, 0.5 * g.ny * g.vox
    g.s_begin, g.s_end, g.t_begin, g.t_end = 0, g.nx, 0, g.ny
    if rng.integers(0, 2):
        a, b = sorted(rng.integers(0, g.nx + 1, 2)); g.s_begin, g.s_end = int(a), int(b)
        a, b = sorted(rng.integers(0, g.ny + 1, 2)); g.t_begin, g.t_end = int(a), int(b)
    view_first = int(rng.integers(0, min(3, g.n_views)))
    sino = rng.random((g.n_views, g.nu), dtype=np.float32) - 0.2
    f_o, img_o = oracle.fbp2(g, sino, view_first=view_first)
    f, img, _ = monte_emu.fbp2(g, sino, view_first=view_first)
    for got, ref, what in ((f, f_o, "filtered"), (img, img_o, "image")):
        scale = float(np.abs(ref).max())
        assert float(np.abs(got.astype(np.float64) - ref).max()) <= REL * scale + 1e-30, (what, seed)


@pytest.mark.parametrize("seed", range(12))
def test_emu_fuzz_partition_invariants(monte_emu, seed):
    """the invariants the multi-GPU plumbing relies on, for random geometries and random cuts: z-slabs, view pieces
    (continuing the stored partial sums) and view-range filtering reproduce the single launch bit for bit; rows
    outside monte_gpu_fdk_slab_rows (set to NaN) are never read by a slab"""
    m = monte_emu
    rng = np.random.default_rng(5000 + seed)
    g = _fdk_case(rng)
    g.full_roi()
    g.mask_r2 = -1
    proj = rng.random((g.n_views, g.nu, g.nv), dtype=np.float32)
    filt = np.full(m.fdk_filtered_shape(g), np.nan, np.float32)
    m.fdk_filter_dev(g, m.Dev(proj), m.Dev(filt))
    whole = np.full((g.nz, g.ny, g.nx), np.nan, np.float32)
    m.fdk_backproject_dev(g, m.Dev(filt), m.Dev(whole))
    assert not np.isnan(whole).any()
    # filter by view ranges
    cuts = sorted(set([0, g.n_views] + [int(x) for x in rng.integers(0, g.n_views + 1, 2)]))
    filt2 = np.full_like(filt, np.nan)
    for a, b in zip(cuts[:-1], cuts[1:]):
        m.fdk_filter_dev(g, m.Dev(proj), m.Dev(filt2), a, b, pad=False)
    m.fdk_pad_dev(g, m.Dev(filt2))
    assert np.array_equal(filt, filt2)
    # z-slabs (+ the rows each slab may read)
    zc = sorted(set([0, g.nz] + [int(x) for x in rng.integers(0, g.nz + 1, 3)]))
    for a, b in zip(zc[:-1], zc[1:]):
        r0, r1 = m.fdk_slab_rows(g, a, b)
        holed = filt.copy()
        rows = holed[: g.n_views * g.nv].reshape(g.n_views, g.nv, -1)
        keep = np.zeros(g.nv, bool)
        keep[r0:r1] = True
        keep[:4] = True
        rows[:, ~keep, :] = np.nan
        slab = np.full((b - a, g.ny, g.nx), np.nan, np.float32)
        m.fdk_backproject_dev(g, m.Dev(holed), m.Dev(slab), a, b)
        assert np.array_equal(slab, whole[a:b]), (a, b, r0, r1)
    # view pieces in ascending order continue the partial sums exactly
    piece = np.full_like(whole, np.nan)
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        m.fdk_backproject_views_dev(g, m.Dev(filt), m.Dev(piece), 0, g.nz, a, b, i > 0)
    assert np.array_equal(piece, whole)


@pytest.mark.parametrize("seed", range(8))
def test_emu_fuzz_fdk_multi_device_pipeline(monte_emu, seed):
    """the chunk-interleaved multi-device reconstruction (2..8 emulated devices, 1..8 view chunks per device, ragged and
    empty chunks, partial ROIs, both weight modes) returns the bits of the one-device call; a one-off run of 60 further
    cases found no failure"""
    import os
    rng = np.random.default_rng(7000 + seed)
    n_views, nu, nv = int(rng.integers(3, 70)), int(rng.integers(12, 60)), int(rng.integers(4, 36))
    n = int(rng.choice([16, 24, 32]))
    nd = int(rng.integers(2, 9))
    chunks = int(rng.integers(1, min(8, 32 // nd) + 1))
    g = _abi.generic_fdk_geom(n_views, nu, nv, n, textbook=bool(rng.integers(0, 2)))
    if rng.integers(0, 3) == 0:
        g.s_begin, g.s_end, g.t_begin, g.t_end = 1, n - 2, 2, n - 1
        g.z_begin, g.z_end = int(rng.integers(0, n // 2)), int(rng.integers(n // 2, n + 1))
    proj = rng.random((n_views, nu, nv), dtype=np.float32)
    keep = {k: os.environ.get(k) for k in ("MONTE_EMU_DEVICES", "MONTE_FDK_MULTI_CHUNKS")}
    os.environ["MONTE_EMU_DEVICES"], os.environ["MONTE_FDK_MULTI_CHUNKS"] = "8", str(chunks)
    try:
        monte_emu.init(0)
        f1, v1, z1, _ = monte_emu.fdk(g, proj, want_zy=True)
        monte_emu.init(list(range(nd)))
        f2, v2, z2, _ = monte_emu.fdk(g, proj, want_zy=True)
    finally:
        monte_emu.init(0)
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert np.array_equal(f1, f2) and np.array_equal(v1, v2) and np.array_equal(z1, z2), (n_views, nu, nv, n, nd, chunks)
