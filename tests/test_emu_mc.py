"""CPU tests of the Monte-Carlo transport kernel, the projector and their host code under SIMT emulation.

monte_b200/csrc/mc.cu and project.cu are compiled by g++ against tests/emu/cuda_runtime.h: the transport
kernel's warps run as 32 fibers each, so its warp votes (__reduce_add_sync / __ballot_sync / __shfl_sync), the
parked-history slots in "shared memory", the unit chaining and the tally flushes are executed exactly as
written.  The bodies of the GPU parity tests are reused (same assertions, same oracle), including the
history-by-history comparison on shared Philox variates.  Test infrastructure only: the product has no CPU
path, speed is not assessed here, and libm replaces the MUFU approximations (fewer fp threshold flips than on
the GPU, never more logic).
"""
import numpy as np
import pytest

import test_mc_gpu as G
from monte_b200 import _abi, scenes


@pytest.mark.parametrize("mode,poly", [(_abi.SOURCE_PENCIL, False), (_abi.SOURCE_CONE, True)])
def test_emu_history_coupled_fates_match_oracle(monte_emu, oracle, mode, poly):
    G.test_history_coupled_fates_match_oracle(monte_emu, oracle, mode, poly)


def test_emu_images_and_counters_match_coupled_oracle(monte_emu, oracle):
    G.test_images_and_counters_match_coupled_oracle(monte_emu, oracle)


def test_emu_energy_integrating_detector_matches_coupled_oracle(monte_emu, oracle):
    G.test_energy_integrating_detector_matches_coupled_oracle(monte_emu, oracle)


def test_emu_three_materials_im_variant_coupled(monte_emu, oracle):
    G.test_three_materials_im_variant_coupled(monte_emu, oracle)


def test_emu_hu_volume_transport_with_present_material_majorant(monte_emu, oracle):
    G.test_hu_volume_transport_with_present_material_majorant(monte_emu, oracle)


def test_emu_scene_update_labels_follows_the_present_materials(monte_emu):
    G.test_scene_update_labels_follows_the_present_materials(monte_emu)


def test_emu_all_air_volume_under_present_majorant(monte_emu):
    G.test_all_air_volume_under_present_majorant(monte_emu)


def test_emu_ring_detector_primary_transmission_kat(monte_emu, oracle):
    G.test_ring_detector_primary_transmission_kat(monte_emu, oracle)


def test_emu_ring_detector_coupled_with_oracle(monte_emu, oracle):
    G.test_ring_detector_coupled_with_oracle(monte_emu, oracle)


def test_emu_edge_cases(monte_emu):
    G.test_edge_cases(monte_emu)


def test_emu_counts_to_map_matches_oracle(monte_emu, oracle):
    G.test_counts_to_map_matches_oracle(monte_emu, oracle)


def test_emu_project_primary_matches_oracle(monte_emu, oracle):
    G.test_project_primary_matches_oracle(monte_emu, oracle)


def test_emu_partition_independence(monte_emu):
    """photon ranges (the multi-GPU split), view ranges and the resident-scene entry point give the tallies and
    counters of the undivided host-buffer call, exactly"""
    m = monte_emu
    g, vol, lab = G.scene(n=33, pitch=1.0, det=17, views=3)
    xs = scenes.make_xs()
    per, seed = 30, 3
    ref0, ref5, st = m.simulate(g, vol, lab, xs, scenes.mono_spectrum(), per, seed)
    sc = m.Scene(g, vol, lab, xs, scenes.mono_spectrum())
    a0, a5 = np.zeros((3, 17, 17), np.int32), np.zeros((3, 17, 17), np.int32)
    stats = np.zeros(16, np.uint64)
    for nr in ((0, 7), (7, 19), (19, 30)):
        sc.simulate_dev(m.Dev(a0), m.Dev(a5), per, seed, views=(0, 2), n_range=nr, d_stats=m.Dev(stats))
    sc.simulate_dev(m.Dev(a0), m.Dev(a5), per, seed, views=(2, 3), d_stats=m.Dev(stats))
    sc.close()
    assert np.array_equal(a0, ref0) and np.array_equal(a5, ref5)
    st2 = m.unpack_stats(stats)
    for k in ("histories", "primaries", "scatter_detected", "absorbed", "interactions", "woodcock_steps"):
        assert st2[k] == st[k], k
    assert st2["sum_e_scatter"] == st["sum_e_scatter"]
    d0, _, _ = m.simulate(g, vol, lab, xs, scenes.mono_spectrum(), per, seed + 1)
    assert not np.array_equal(d0, ref0)


@pytest.mark.parametrize("which", ["31", "33", "36", "44", "45", "46"])
def test_emu_kernel_variants_give_identical_tallies(which, oracle):
    """MONTE_MC_KERNEL selects K = 1..6 parked histories per lane and the one- / three-slot step visits: every
    history consumes its own counter-based variates, so all variants must produce the default kernel's tallies
    bit for bit.  The variant is latched at the first launch of a process, hence the subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import os, sys, importlib.util
import numpy as np
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
spec = importlib.util.spec_from_file_location("monte_emu_build", os.path.join(root, "tests", "emu", "build.py"))
eb = importlib.util.module_from_spec(spec); spec.loader.exec_module(eb)
m = eb.api(); m.init(0)
import test_mc_gpu as G
from monte_b200 import scenes
g, vol, lab = G.scene(n=33, pitch=1.0, det=17, views=2)
im0, im5, st = m.simulate(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum(140.0), 25, 9)
np.save(sys.argv[2], np.stack([im0, im5]))
print(st["histories"], st["woodcock_steps"], st["interactions"])
'''
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    with tempfile.TemporaryDirectory() as td:
        for w in ("35", which):
            env = dict(os.environ, MONTE_MC_KERNEL=w)
            path = os.path.join(td, "im_%s.npy" % w)
            r = subprocess.run([sys.executable, "-c", code, root, path], env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            outs[w] = (np.load(path), r.stdout.strip())
    assert np.array_equal(outs["35"][0], outs[which][0])
    assert outs["35"][1] == outs[which][1]


# ---- Rayleigh form-factor deflection (coherent_mode = FORMFACTOR): the GPU test bodies of
# tests/test_rayleigh_gpu.py on the emulated kernel
import test_rayleigh_gpu as R   # noqa: E402


@pytest.mark.parametrize("keV,poly", [(60.0, False), (30.0, False), (0.0, True)])
def test_emu_formfactor_history_coupled_fates_match_oracle(monte_emu, oracle, keV, poly):
    R.test_formfactor_history_coupled_fates_match_oracle(monte_emu, oracle, keV, poly)


def test_emu_formfactor_images_counters_and_difference_from_forward_mode(monte_emu, oracle):
    R.test_formfactor_images_counters_and_difference_from_forward_mode(monte_emu, oracle)


def test_emu_formfactor_mode_needs_tables(monte_emu):
    R.test_formfactor_mode_needs_tables(monte_emu)


# ---- two-level Woodcock majorant (tracking_mode = CLEARANCE): the GPU test bodies of tests/test_tracking_gpu.py
import test_tracking_gpu as CL   # noqa: E402


@pytest.mark.parametrize("cell_log2,poly,rayleigh", [(0, True, False), (1, True, False), (2, False, False), (1, True, True)])
def test_emu_clearance_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly, rayleigh):
    CL.test_clearance_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly, rayleigh)


def test_emu_clearance_counters_and_fewer_steps_than_the_reference_loop(monte_emu, oracle):
    CL.test_clearance_counters_and_fewer_steps_than_the_reference_loop(monte_emu, oracle)


def test_emu_clearance_partition_and_label_update(monte_emu):
    CL.test_clearance_partition_and_label_update(monte_emu)
    # resident scene: new labels through monte_gpu_scene_update_labels rebuild the clearance grid
    m = monte_emu
    g, vol, lab = G.scene(n=33, pitch=1.0, det=9, views=1)
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 1
    xs, spec = scenes.make_xs(), scenes.mono_spectrum(50.0)
    lab2 = np.ascontiguousarray(lab[:, :, ::-1])                    # mirrored phantom: rods elsewhere
    ref0, ref5, _ = m.simulate(g, vol, lab2, xs, spec, 30, 8)
    sc = m.Scene(g, vol, lab, xs, spec)
    sc.update_labels(lab2)
    a0, a5 = np.zeros((1, 9, 9), np.int32), np.zeros((1, 9, 9), np.int32)
    sc.simulate_dev(m.Dev(a0), m.Dev(a5), 30, 8)
    sc.close()
    assert np.array_equal(a0, ref0) and np.array_equal(a5, ref5)


def test_emu_clearance_grid_cache_follows_the_label_content(monte_emu):
    """the host-buffer call keeps the clearance grid of the last labels (keyed by a content hash): the same buffer
    with new content must rebuild it"""
    m = monte_emu
    g, vol, lab = G.scene(n=33, pitch=1.0, det=9, views=1)
    vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 0
    xs, spec = scenes.make_xs(), scenes.mono_spectrum(50.0)
    buf = lab.copy()
    a0, a5, _ = m.simulate(g, vol, buf, xs, spec, 30, 8)
    b0, b5, _ = m.simulate(g, vol, buf, xs, spec, 30, 8)                 # cached grid
    assert np.array_equal(a0, b0) and np.array_equal(a5, b5)
    buf[:] = np.roll(lab, 5, axis=2)                                     # same buffer, the rods moved
    c0, c5, _ = m.simulate(g, vol, buf, xs, spec, 30, 8)
    sc = m.Scene(g, vol, buf.copy(), xs, spec)                           # a fresh scene builds its own grid
    d0, d5 = np.zeros((1, 9, 9), np.int32), np.zeros((1, 9, 9), np.int32)
    sc.simulate_dev(m.Dev(d0), m.Dev(d5), 30, 8)
    sc.close()
    assert np.array_equal(c0, d0) and np.array_equal(c5, d5)
    assert not np.array_equal(c5, a5)


def test_emu_auto_tracking_equals_the_mode_it_resolves_to(monte_emu):
    m = monte_emu
    g, vol, lab = G.scene(n=33, pitch=1.0, det=9, views=1)
    xs = scenes.make_xs()
    for spec in (scenes.mono_spectrum(140.0), scenes.mono_spectrum(50.0)):
        mode, cl, ratio = m.resolve_tracking(xs, spec)
        vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_AUTO, 0
        a0, a5, sa = m.simulate(g, vol, lab, xs, spec, 30, 4)
        vol.tracking_mode, vol.clearance_cell_log2 = mode, cl
        b0, b5, sb = m.simulate(g, vol, lab, xs, spec, 30, 4)
        assert np.array_equal(a0, b0) and np.array_equal(a5, b5) and sa["woodcock_steps"] == sb["woodcock_steps"]
    assert m.resolve_tracking(xs, scenes.mono_spectrum(140.0))[0] == _abi.TRACK_GLOBAL
    assert m.resolve_tracking(xs, scenes.mono_spectrum(50.0))[0] == _abi.TRACK_DIRECTIONAL


@pytest.mark.parametrize("sched", ["reverse", "random"])
def test_emu_results_do_not_depend_on_the_thread_schedule(sched, tmp_path):
    """racecheck-lite: the emulator resumes runnable threads in reverse or random order (MONTE_EMU_SCHED) -- every
    kernel with complete barriers must give bit-identical results; MC tallies (integer atomics), the direct and FFT
    filters (shared-memory tiles and exchange buffers), backprojection, transpose, projector"""
    import os
    import subprocess
    import sys
    code = r'''
import os, sys, importlib.util
import numpy as np
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
spec = importlib.util.spec_from_file_location("monte_emu_build", os.path.join(root, "tests", "emu", "build.py"))
eb = importlib.util.module_from_spec(spec); spec.loader.exec_module(eb)
m = eb.api(); m.init(0)
import test_mc_gpu as G
from monte_b200 import _abi, scenes
g, vol, lab = G.scene(n=33, pitch=1.0, det=17, views=2)
vol.tracking_mode, vol.clearance_cell_log2 = 1, 1
g.coherent_mode = 1
xs = scenes.add_formfactors(scenes.make_xs())
im0, im5, st = m.simulate(g, vol, lab, xs, scenes.mono_spectrum(60.0), 25, 9)
fg = _abi.generic_fdk_geom(24, 65, 33, 32)
f, v, zy, _ = m.fdk(fg, np.random.default_rng(1).random((24, 65, 33), dtype=np.float32), want_zy=True)
gw = _abi.generic_fdk_geom(2, 300, 5, 8)
fw = m.fdk(gw, np.random.default_rng(2).random((2, 300, 5), dtype=np.float32))[0]
pm = m.project_primary(g, vol, lab, xs, 60.0)
np.savez(sys.argv[2], im0=im0, im5=im5, f=f, v=v, zy=zy, fw=fw, pm=pm, steps=st["woodcock_steps"])
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("forward", sched):
        path = os.path.join(str(tmp_path), mode + ".npz")
        r = subprocess.run([sys.executable, "-c", code, root, path], env=dict(os.environ, MONTE_EMU_SCHED=mode),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(path))
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_emu_resident_projector_feeds_fdk_like_the_host_path(monte_emu, oracle):
    CL.test_resident_projector_feeds_fdk_like_the_host_path(monte_emu, oracle)


@pytest.mark.parametrize("cell_log2,poly", [(0, True), (1, False)])
def test_emu_adaptive_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly):
    CL.test_adaptive_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly)


def test_emu_adaptive_never_needs_more_steps_than_the_reference_loop(monte_emu):
    CL.test_adaptive_never_needs_more_steps_than_the_reference_loop(monte_emu)


def test_emu_cbct_mc_driver_end_to_end(monte_emu, tmp_path):
    """the C++ driver main() of monte_b200/host/cbct_mc.cpp linked against the emulated library: argument parsing,
    CSV tables with BOM/CRLF, raw label input, the reference's four output files -- with and without the optional
    transport modes; the count images equal the same run through the Python binding"""
    import os
    import subprocess
    import test_drivers as D
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emu_dir = os.path.join(root, "tests", "emu", "_build")
    exe = os.path.join(str(tmp_path), "cbct_mc_emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "monte_b200", "host", "cbct_mc.cpp"), "-o", exe,
                           "-L" + emu_dir, "-lmonte_gpu_emu", "-Wl,-rpath," + emu_dir])
    d = str(tmp_path)
    h2o, ca = scenes.load_tables()
    D._write_csv(os.path.join(d, "xcom2.csv"), h2o)
    D._write_csv(os.path.join(d, "Ca.csv"), ca)
    lab = scenes.cylinder_phantom(33, 1.0)
    lab.tofile(os.path.join(d, "cyl.raw"))
    for tag, extra in (("a", []), ("b", ["1", "1"])):                 # b: rayleigh = 1, clearance cells of 2 voxels
        out = subprocess.run([exe, "cyl.raw", "33", "1.0", "xcom2.csv", "Ca.csv", "9", str(32.5 / 9), "2", "40", "3", tag] + extra,
                             cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
        assert "histories" in out
        p0 = np.fromfile(os.path.join(d, "proj_%s0.raw" % tag), np.int32).reshape(2, 9, 9)
        p5 = np.fromfile(os.path.join(d, "proj_%s5.raw" % tag), np.int32).reshape(2, 9, 9)
        m0 = np.fromfile(os.path.join(d, "map_%s0.raw" % tag), np.float32).reshape(2, 9, 9)
        assert (p5 >= p0).all() and p0.max() <= 40
        assert np.abs(m0 - (-np.log(np.clip(p0, 1, 40).astype(np.float64)) + np.log(40.0))).max() < 1e-5
        # the same run through the Python binding (the driver's geometry: 180 degrees apart, untight clip box, BOM-quirk tables)
        g = scenes.mc_geom(9, 32.5 / 9, n_views=2)
        g.angle_step_deg = 180.0
        vol = scenes.volume_for(lab, 1.0, tight=False)
        xs = scenes.make_xs(quirk_bom=True)
        if extra:
            g.coherent_mode = _abi.COHERENT_FORMFACTOR
            scenes.add_formfactors(xs)
            vol.tracking_mode, vol.clearance_cell_log2 = _abi.TRACK_CLEARANCE, 1
        r0, r5, _ = monte_emu.simulate(g, vol, lab, xs, scenes.mono_spectrum(140.0), 40, 3)
        assert np.array_equal(r0, p0) and np.array_equal(r5, p5), tag
    # --hu: a CT volume in Hounsfield units goes in; the driver segments it (monte_ctnum_segment, default classes) and
    # tracks with the majorant of the classes present -- equal to the same chain through the Python binding
    hu = scenes.hu_head_phantom(33, 0.6)
    hu.tofile(os.path.join(d, "head_hu.raw"))
    out = subprocess.run([exe, "head_hu.raw", "33", "0.6", "xcom2.csv", "Ca.csv", "9", str(32.5 / 9), "2", "40", "3", "h", "--hu", "--kev", "60"],
                         cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    assert "segmented into 8 classes, present mask 0x4f" in out
    p0 = np.fromfile(os.path.join(d, "proj_h0.raw"), np.int32).reshape(2, 9, 9)
    p5 = np.fromfile(os.path.join(d, "proj_h5.raw"), np.int32).reshape(2, 9, 9)
    lab_h, xs_h, _, present = monte_emu.ctnum_segment(hu, monte_emu.hu_classes_default(True), scenes.make_xs(quirk_bom=True), 60.0)
    g = scenes.mc_geom(9, 32.5 / 9, n_views=2)
    g.angle_step_deg = 180.0
    vol = scenes.volume_for(lab_h, 0.6, tight=False)
    vol.majorant_mode = _abi.MAJORANT_PRESENT
    r0, r5, _ = monte_emu.simulate(g, vol, lab_h, xs_h, scenes.mono_spectrum(60.0), 40, 3)
    assert np.array_equal(r0, p0) and np.array_equal(r5, p5) and 0 < p0.sum() < 2 * 81 * 40
    # --ring R: the ring detector (source at the origin, 12 angular bins, one axial bin)
    out = subprocess.run([exe, "cyl.raw", "33", "1.0", "xcom2.csv", "Ca.csv", "12", "2.0", "1", "50", "3", "r", "--ring", "25"],
                         cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    p0 = np.fromfile(os.path.join(d, "proj_r0.raw"), np.int32).reshape(1, 12, 1)
    p5 = np.fromfile(os.path.join(d, "proj_r5.raw"), np.int32).reshape(1, 12, 1)
    g = scenes.ring_geom(12, 1, 2.0, 25.0)
    g.angle_step_deg = 360.0
    vol = scenes.volume_for(lab, 1.0, tight=False)
    r0, r5, _ = monte_emu.simulate(g, vol, lab, scenes.make_xs(quirk_bom=True), scenes.mono_spectrum(140.0), 50, 3)
    assert np.array_equal(r0, p0) and np.array_equal(r5, p5) and 0 < p0.sum() < 12 * 50


@pytest.mark.parametrize("cell_log2,poly,rayleigh", [(0, True, False), (1, False, False), (1, True, True)])
def test_emu_directional_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly, rayleigh):
    CL.test_directional_history_coupled_fates_match_oracle(monte_emu, oracle, cell_log2, poly, rayleigh)
