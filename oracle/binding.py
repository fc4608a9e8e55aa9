"""ctypes binding of oracle/liboracle.so and runners for the oracle/_ref binaries.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under monte_b200/ imports this module.
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

from monte_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

_lib = None


def build(ref=True):
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(ref=False)
        _lib = C.CDLL(LIB)
        _lib.oracle_fdk.restype = C.c_int
    return _lib


def have_ref(name):
    return os.path.exists(os.path.join(REF_DIR, name))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ---------------------------------------------------------------- FDK
def fdk(g, proj, want_zy=False):
    """Oracle FDK.  proj [views][nu][nv] float32 -> (filtered [views][nv][nu], vol_xy, vol_zy|None)."""
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    assert proj.shape == (g.n_views, g.nu, g.nv)
    filt = np.empty((g.n_views, g.nv, g.nu), np.float32)
    vol = np.empty((g.nz, g.ny, g.nx), np.float32)
    vzy = np.empty((g.nx, g.ny, g.nz), np.float32) if want_zy else None
    rc = lib().oracle_fdk(C.byref(g), _fp(proj), _fp(filt), _fp(vol), _fp(vzy) if want_zy else None)
    assert rc == 0
    return filt, vol, vzy


def fdk_filter(g, proj):
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    mw = np.empty_like(proj)
    filt = np.empty((g.n_views, g.nv, g.nu), np.float32)
    lib().oracle_fdk_weight(C.byref(g), _fp(proj), _fp(mw))
    lib().oracle_fdk_filter(C.byref(g), _fp(mw), _fp(filt))
    return filt


def fdk_backproject(g, filt):
    filt = np.ascontiguousarray(filt, dtype=np.float32)
    vol = np.zeros((g.nz, g.ny, g.nx), np.float32)
    lib().oracle_fdk_backproject(C.byref(g), _fp(filt), _fp(vol), None)
    return vol


def fbp2(g, sino, view_first=1):
    sino = np.ascontiguousarray(sino, dtype=np.float32)
    filt = np.empty((g.n_views, g.nu), np.float32)
    img = np.empty((g.ny, g.nx), np.float32)
    lib().oracle_fbp2(C.byref(g), C.c_int(view_first), _fp(sino), _fp(filt), _fp(img))
    return filt, img


# ---------------------------------------------------------------- unmodified reference binaries
def _run_ref(name, inputs, outputs, timeout=900):
    """Run oracle/_ref/<name> in a scratch dir holding `inputs` {filename: bytes}; return
    ({filename: bytes} for `outputs`, stdout)."""
    exe = os.path.join(REF_DIR, name)
    tmp = tempfile.mkdtemp(prefix="monte_ref_")
    try:
        for fn, data in inputs.items():
            with open(os.path.join(tmp, fn), "wb") as f:
                f.write(data)
        p = subprocess.run([exe], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
        res = {}
        for fn in outputs:
            with open(os.path.join(tmp, fn), "rb") as f:
                res[fn] = f.read()
        return res, p.stdout.decode("utf-8", "replace")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_bp3d20(proj):
    """recon/bp3d20.cpp as shipped: proj [360][65][65] -> (filtered [360][65][65], vol_xy 256^3, vol_zy)."""
    out, _ = _run_ref("bp3d20", {"mapg_0_20000h2oCa.raw": np.ascontiguousarray(proj, np.float32).tobytes()},
                      ["CBCTrecon\\map2mh20kyu2.raw", "CBCTrecon\\xy2mh20kyu2.raw", "CBCTrecon\\zy2mh20kyu2.raw"])
    f = np.frombuffer(out["CBCTrecon\\map2mh20kyu2.raw"], np.float32).reshape(360, 65, 65)
    xy = np.frombuffer(out["CBCTrecon\\xy2mh20kyu2.raw"], np.float32).reshape(256, 256, 256)
    zy = np.frombuffer(out["CBCTrecon\\zy2mh20kyu2.raw"], np.float32).reshape(256, 256, 256)
    return f, xy, zy


def ref_bp3d20_325(proj):
    out, _ = _run_ref("bp3d20_325", {"mapg325_0_2tca.raw": np.ascontiguousarray(proj, np.float32).tobytes()},
                      ["CBCTrecon\\mapca0_2t.raw", "CBCTrecon\\xyca0_2t.raw", "CBCTrecon\\zyca0_2t.raw"],
                      timeout=3600)
    f = np.frombuffer(out["CBCTrecon\\mapca0_2t.raw"], np.float32).reshape(360, 325, 325)
    xy = np.frombuffer(out["CBCTrecon\\xyca0_2t.raw"], np.float32).reshape(256, 256, 256)
    zy = np.frombuffer(out["CBCTrecon\\zyca0_2t.raw"], np.float32).reshape(256, 256, 256)
    return f, xy, zy


def ref_fbp2(sino):
    out, _ = _run_ref("fbp2", {"map5_20_2e5.raw": np.ascontiguousarray(sino, np.float32).tobytes()},
                      ["projection_test\\outproj2e5.raw", "projection_test\\testxyo2e5.raw"])
    f = np.frombuffer(out["projection_test\\outproj2e5.raw"], np.float32).reshape(360, 65)
    img = np.frombuffer(out["projection_test\\testxyo2e5.raw"], np.float32).reshape(256, 256)
    return f, img
