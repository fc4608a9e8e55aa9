"""TEST INFRASTRUCTURE: build tests/emu/_build/libmonte_gpu_emu.so — monte_b200/csrc/*.cu compiled by g++
against tests/emu/cuda_runtime.h (SIMT emulation on fibers) — and hand out a copy of monte_b200.api bound
to it.  Used only by tests/test_emu_*.py to exercise kernel and host logic on the CPU; the product
(monte_b200/lib/libmonte_gpu.so, nvcc, sm_100a) never contains or loads any of this.
"""
import importlib.util
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "monte_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmonte_gpu_emu.so")

CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I" + HERE]
CU = ["common.cu", "fdk.cu", "fbp2.cu", "mc.cu", "project.cu"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, asan=False, ubsan=False):
    """asan=True: a second copy under _build_asan/ compiled with -fsanitize=address (device buffers are heap
    blocks there, so an out-of-bounds access of a kernel or of the host code is reported like by compute-sanitizer
    memcheck); load it in a process started with LD_PRELOAD=libasan (tests/emu/asan_check.py)."""
    out_dir = OUT + "_asan" if asan else OUT + "_ubsan" if ubsan else OUT
    lib_path = os.path.join(out_dir, "libmonte_gpu_emu.so")
    san = "address" if asan else "undefined" if ubsan else None       # ubsan=True: same idea with -fsanitize=undefined
    flags = CXXFLAGS + (["-fsanitize=" + san, "-fno-omit-frame-pointer", "-O1"] if san else [])
    os.makedirs(out_dir, exist_ok=True)
    import fcntl
    with open(os.path.join(out_dir, ".lock"), "w") as lock:           # several pytest-xdist workers may get here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build_locked(force, out_dir, lib_path, san, flags)


def _build_locked(force, out_dir, lib_path, san, flags):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hdrs += [os.path.join(ROOT, "include", "monte_gpu.h"), os.path.join(HERE, "cuda_runtime.h")]
    jobs, objs = [], []
    for src, lang in [(os.path.join(CSRC, f), ["-x", "c++"]) for f in CU] + \
                     [(os.path.join(CSRC, "host_helpers.cpp"), []), (os.path.join(HERE, "emu_runtime.cpp"), [])]:
        obj = os.path.join(out_dir, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append((src, subprocess.Popen(["g++"] + flags + lang + ["-c", src, "-o", obj],
                                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in jobs:
        out = p.communicate()[0]
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("g++ (emulation build) failed on %s" % src)
    if force or _newer(lib_path, objs):
        subprocess.check_call(["g++", "-shared"] + (["-fsanitize=" + san] if san else []) + ["-o", lib_path] + objs)
    return lib_path


class Dev:
    """a "device" buffer of the emulation: host memory behind the .data_ptr() the *_dev wrappers of api.py ask for"""

    def __init__(self, a):
        import numpy as np
        self.a = a
        assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]

    def data_ptr(self):
        return self.a.ctypes.data

    def __getitem__(self, k):
        return Dev(self.a[k])


_api = None


def api(asan=False, ubsan=False):
    """A private copy of the monte_b200.api module whose entry points call the emulation library."""
    global _api
    if _api is None:
        lib_path = build(asan=asan, ubsan=ubsan)
        spec = importlib.util.spec_from_file_location("monte_b200._api_emu", os.path.join(ROOT, "monte_b200", "api.py"),
                                                      submodule_search_locations=None)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "monte_b200"
        spec.loader.exec_module(mod)
        mod._lib = mod._bind(lib_path)
        mod._stream_ptr = lambda stream=None: None       # "device" pointers are host pointers, streams run in issue order
        mod.Dev = Dev
        assert not mod._lib._monte_missing, mod._lib._monte_missing
        _api = mod
    return _api


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
