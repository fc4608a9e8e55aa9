"""TEST INFRASTRUCTURE: run the smoke path and a cross-section of the parity-test bodies on the emulated library
built with AddressSanitizer.  Device buffers are heap blocks in the emulation, so any out-of-bounds read or write
of a kernel (labels, filtered rows, volumes, tallies, tables, clearance grid) or of the C-ABI host code aborts with
an ASan report -- the CPU counterpart of `compute-sanitizer --tool memcheck` (profiles/README.md).  Start with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emu/asan_check.py
(tests/test_emu_fdk.py::test_emu_address_sanitizer_clean does).  Dynamic shared memory is a heap block of exactly
the size the launch asked for, static __shared__ arrays are instrumented statics, so overruns of either are
reported too; racecheck has no counterpart (fibers run one at a time).
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    spec = importlib.util.spec_from_file_location("monte_emu_build", os.path.join(HERE, "build.py"))
    eb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eb)
    ubsan = "--ubsan" in sys.argv      # LD_PRELOAD=$(gcc -print-file-name=libubsan.so) python tests/emu/asan_check.py --ubsan
    m = eb.api(asan=not ubsan, ubsan=ubsan)
    import __graft_entry__ as ge
    ge._smoke(m)
    m.init(0)
    from monte_b200 import _abi
    from oracle import binding as ob
    import test_fdk_gpu as F
    import test_mc_gpu as G
    import test_rayleigh_gpu as R
    import test_tracking_gpu as T
    F.test_generic_geometry_against_oracle(m, ob, 96, 40, 40, 30, False)        # ragged sizes
    F.test_partial_roi_and_mask(m, ob)
    F.test_empty_roi_gives_zero_volume(m)
    F.test_fbp2_against_reference_golden(m)
    G.test_edge_cases(m)
    G.test_three_materials_im_variant_coupled(m, ob)
    G.test_project_primary_matches_oracle(m, ob)
    T.test_clearance_history_coupled_fates_match_oracle(m, ob, 0, True, True)
    T.test_clearance_partition_and_label_update(m)
    R.test_formfactor_images_counters_and_difference_from_forward_mode(m, ob)
    g = _abi.generic_fdk_geom(64, 48, 40, 128)                                   # 8 view chunks, 4 z-slabs, transposed copy
    g.s_begin, g.s_end, g.t_begin, g.t_end = 56, 72, 50, 66
    m.fdk(g, F.rand(21, (64, 48, 40)), want_zy=True)
    gw = _abi.generic_fdk_geom(2, 1100, 3, 16)                                   # FFT filter, L = 4096, lone last column
    m.fdk(gw, F.rand(1, (2, 1100, 3)))
    m.shutdown()
    print("ASAN RUN COMPLETE")


if __name__ == "__main__":
    main()
