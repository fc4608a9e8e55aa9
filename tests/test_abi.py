"""CPU: the C-ABI library loads without a GPU and exports every symbol include/monte_gpu.h declares;
struct layouts of the ctypes mirror match the header (compiled probe); the product never imports
the oracle; entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from monte_b200 import _abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "monte_gpu.h")


@pytest.fixture(scope="module")
def lib():
    from monte_b200 import build
    build.build_lib()
    return api.load()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(monte_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libmonte_gpu.so does not export %s" % n
    assert lib._monte_missing == []
    assert set(lib._monte_symbols) == set(names), set(lib._monte_symbols) ^ set(names)
    assert lib.monte_gpu_abi_version() == 7


def test_struct_layouts_match_the_header():
    probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "monte_gpu.h"
int main(void){
 printf("%zu %zu %zu ", sizeof(monte_hu_class), offsetof(monte_hu_class, density), offsetof(monte_mc_volume, majorant_mode));
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(monte_fdk_geom), sizeof(monte_fdk_stats),
   sizeof(monte_mc_xs), sizeof(monte_mc_volume), sizeof(monte_mc_geom), sizeof(monte_mc_spectrum),
   sizeof(monte_mc_stats), offsetof(monte_fdk_geom, mask_r2), offsetof(monte_fdk_geom, coord_mode),
   offsetof(monte_mc_geom, max_scatter), offsetof(monte_mc_stats, sum_e_primary));
 return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "p.c"), "w") as f:
            f.write(probe)
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    want = [C.sizeof(_abi.HuClass), _abi.HuClass.density.offset, _abi.McVolume.majorant_mode.offset,
            C.sizeof(_abi.FdkGeom), C.sizeof(_abi.FdkStats), C.sizeof(_abi.McXs), C.sizeof(_abi.McVolume),
            C.sizeof(_abi.McGeom), C.sizeof(_abi.McSpectrum), C.sizeof(_abi.McStats),
            _abi.FdkGeom.mask_r2.offset, _abi.FdkGeom.coord_mode.offset, _abi.McGeom.max_scatter.offset,
            _abi.McStats.sum_e_primary.offset]
    assert got == want


def test_presets_equal_the_shipped_literals(lib):
    for fn, py in ((lib.monte_fdk_geom_bp3d20, _abi.bp3d20_geom()), (lib.monte_fdk_geom_bp3d20_325, _abi.bp3d20_325_geom()),
                   (lib.monte_fdk_geom_fbp2, _abi.fbp2_geom())):
        g = _abi.FdkGeom()
        fn(C.byref(g))
        for name, _ in _abi.FdkGeom._fields_:
            assert getattr(g, name) == getattr(py, name), name


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "monte_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), fn      # comments may cite it
                assert "liboracle" not in text, fn
                # the one dlopen of the product binds NCCL for the multi-device mode (csrc/common.cu); nothing else is loaded at run time
                if "dlopen(" in text:
                    assert fn == "common.cu" and set(re.findall(r'"(lib[^"]*\.so[^"]*)"', text)) == {"libnccl.so.2", "libnccl.so"}, fn


def test_product_never_contains_or_loads_the_emulation():
    """tests/emu (the SIMT emulation the CPU suite uses) is test infrastructure: libmonte_gpu.so is built by nvcc
    without MONTE_EMU, exports nothing of it, and nothing under monte_b200/ points at the emulation library"""
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "monte_emu" not in syms
    for dirpath, _, files in os.walk(os.path.join(ROOT, "monte_b200")):
        for fn in files:
            if fn.endswith((".py", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "libmonte_gpu_emu" not in text and "MONTE_EMU" not in text, fn
    build_py = open(os.path.join(ROOT, "monte_b200", "build.py")).read()
    assert "tests" not in build_py.replace("tests/emu/build.py", "")


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.MonteError, match="no CUDA device|CUDA"):
        api.init(0)
    g = _abi.generic_fdk_geom(4, 16, 16, 8)
    import numpy as np
    with pytest.raises(api.MonteError, match="monte_gpu_init"):
        api.fdk(g, np.zeros((4, 16, 16), np.float32))


def test_host_helpers(lib):
    import numpy as np
    a = np.zeros((65, 65), np.uint8)
    lib.monte_make_fantom(C.c_void_p(a.ctypes.data), 65, 32, 46, 100)       # make_fantom.cpp:10-19
    jj, kk = np.ogrid[:65, :65]
    assert np.array_equal(a, (((jj - 32) ** 2 + (kk - 46) ** 2) <= 100).astype(np.uint8)) and a.sum() == 317
    s = np.zeros((13, 9, 9), np.uint8)
    lib.monte_make_sphere(C.c_void_p(s.ctypes.data), 9, 9, 13, 4, 4, 6, 9)
    assert s[6, 4, 4] == 1 and s[6, 4, 7] == 1 and s[6, 4, 8] == 0 and s.sum() == 123
    # CSV loader against the packed tables (written back out in the reference's layout)
    from monte_b200 import scenes
    h2o, _ = scenes.load_tables()
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.csv")
        with open(p, "wb") as f:
            f.write(b"\xef\xbb\xbf" + b"\r\n".join(b"%.10g,%.10g,%.10g,%.10g" % tuple(h2o[:, k]) for k in range(1, 201)) + b"\r\n")
        xs = _abi.McXs()
        assert lib.monte_xs_load_csv(p.encode(), 0, 1.0, 1, C.byref(xs)) == 0
        assert xs.n_materials == 1 and abs(xs.total[0][140] - 0.1538092) < 1e-7
        assert abs(xs.coh[0][1] - 1.372) < 1e-6 and xs.total[0][0] == xs.total[0][1]
        assert lib.monte_xs_load_csv(b"/nonexistent.csv", 0, 1.0, 0, C.byref(xs)) == -6
        assert b"cannot open" in lib.monte_gpu_last_error()
    hu = np.array([-1000, 0, 1000, 3000], np.float32)
    mu = np.zeros(4, np.float32)
    lab = np.zeros(4, np.uint8)
    xs2 = scenes.make_xs()
    assert lib.monte_ctnum_to_mu(C.c_void_p(hu.ctypes.data), 4, C.byref(xs2), 140.0, -500.0, 700.0,
                                 C.c_void_p(mu.ctypes.data), C.c_void_p(lab.ctypes.data)) == 0
    assert mu[0] == 0 and abs(mu[1] - 0.1538092) < 1e-6 and abs(mu[2] - 2 * 0.1538092) < 1e-6
    assert list(lab) == [0, 1, 2, 2]
    # per-keV majorant over the materials present (CBCT_real325im.cu:866 restricted to the volume's labels)
    mm = np.zeros(201, np.float32)
    assert lib.monte_xs_majorant(C.byref(xs2), lab.ctypes.data, 4, mm.ctypes.data) == 0
    assert abs(mm[140] - 0.17730 * 1.55) < 1e-4                    # calcium sets it
    water_only = np.array([0, 1, 1, 0], np.uint8)
    assert lib.monte_xs_majorant(C.byref(xs2), water_only.ctypes.data, 4, mm.ctypes.data) == 0
    assert abs(mm[140] - 0.1538092) < 1e-6
    assert lib.monte_xs_majorant(C.byref(xs2), None, 0, mm.ctypes.data) == 0 and abs(mm[60] - max(0.20585, xs2.total[1][60] * 1.55)) < 1e-4
    assert lib.monte_xs_majorant(None, None, 0, mm.ctypes.data) == -1


def test_fft_core_on_the_host(tmp_path):
    """monte_b200/csrc/fft_core.cuh is plain C++: its per-thread FFT phases, driven by a sequential loop
    over the threads, reproduce the direct Ram-Lak convolution of recon/bp3d20.cpp:63-73 (all three
    transform lengths, full and partial rows)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fft_core_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "monte_b200", "csrc"),
                    os.path.join(root, "tests", "fft_core_host.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("rel_err") == 6


@pytest.mark.parametrize("L", [1024, 2048, 4096])
def test_fft_exchange_swizzle_is_bank_conflict_free(L):
    """fft_pad (csrc/fft_core.cuh) must be a bijection and put the 16 eight-byte elements a half-warp
    touches in one access on 16 distinct bank pairs, for the loads and the stores of every pass."""
    def pad(i):
        m = i >> 4
        return i ^ ((m & 7) | ((m & 4) << 1))
    assert sorted(pad(i) for i in range(L)) == list(range(L))
    T, r_last = L // 8, L // 512
    patterns = [lambda j, r: j + r * T,                                     # loads of the radix-8 passes
                lambda j, q: 8 * j + q,                                     # stores, Ns = 1
                lambda j, q: (j // 8) * 64 + j % 8 + 8 * q,                 # stores, Ns = 8
                lambda j, q: (j // 64) * 512 + j % 64 + 64 * q,             # stores, Ns = 64
                lambda j, q: j + (q % r_last) * (L // r_last),              # loads of the last pass
                lambda j, q: j + (q % r_last) * 512]                        # stores of the last pass
    for f in patterns:
        for hw in range(T // 16):
            for q in range(8):
                banks = {pad(f(16 * hw + lane, q)) % 16 for lane in range(16)}
                assert len(banks) == 16
