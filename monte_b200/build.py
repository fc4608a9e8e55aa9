"""Build libmonte_gpu.so (CUDA, sm_100a only) and the C++ drivers, in-tree.

    python -m monte_b200.build            # library + drivers
nvcc cross-compiles without a GPU.  The .so lands in monte_b200/lib/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libmonte_gpu.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]
CU_SOURCES = ["common.cu", "fdk.cu", "fbp2.cu", "mc.cu", "project.cu"]
# mc.cu: flush-to-zero arithmetic.  Its MUFU calls (lg2 of a uniform >= 2^-24, rcp / rsq of O(1) quantities) never see or
# produce a denormal that matters, but without the flag every one of them carries a 3-instruction denormal guard
# (FSETP |x| < 2^-126, FMUL 2^24, FADD -24): 14 guards in the transport kernel.  Results for normal operands are identical.
EXTRA_FLAGS = {"mc.cu": ["-ftz=true"]}
CPP_SOURCES = ["host_helpers.cpp"]
DRIVERS = ["make_fantom", "ctnum_to_mu", "cbct_mc", "cbct_fdk"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "monte_gpu.h"))
    objs = []
    log = []
    for src in CU_SOURCES + CPP_SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            raise FileNotFoundError(sp)
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        if force or _newer(obj, [sp, os.path.abspath(__file__)] + headers):
            cmd = [NVCC] + ARCH + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", sp, "-o", obj]
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            log.append(p.stdout)
            if verbose or p.returncode != 0:
                sys.stderr.write(p.stdout)
            if p.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    with open(os.path.join(objdir, "ptxas.log"), "a") as f:
        f.write("\n".join(log))
    return LIB


def build_drivers(force=False):
    os.makedirs(BINDIR, exist_ok=True)
    out = []
    for d in DRIVERS:
        src = os.path.join(HOST, d + ".cpp")
        if not os.path.exists(src):
            continue
        exe = os.path.join(BINDIR, d)
        if force or _newer(exe, [src, LIB, os.path.join(ROOT, "include", "monte_gpu.h")]):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                                   "-L" + LIBDIR, "-lmonte_gpu", "-Wl,-rpath,$ORIGIN/../lib",
                                   "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
        out.append(exe)
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_lib(force=force, verbose="-v" in sys.argv))
    for e in build_drivers(force=force):
        print(e)
