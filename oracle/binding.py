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


def set_threads(n=0):
    """OpenMP threads of the oracle's loops: n > 0 sets them (overriding OMP_NUM_THREADS); returns the number in force"""
    l = lib()
    l.oracle_set_threads.restype = C.c_int
    return int(l.oracle_set_threads(C.c_int(int(n))))


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


# ---------------------------------------------------------------- Monte Carlo
Q_REAL2 = (1 << 15) - 1
RNG_MT, RNG_PHILOX = 0, 1


class OracleTables(C.Structure):
    _fields_ = [("n_materials", C.c_int32),
                ("density", C.c_double * _abi.MAX_MATERIALS),
                ("coh", (C.c_double * _abi.TABLE_ROWS) * _abi.MAX_MATERIALS),
                ("compt", (C.c_double * _abi.TABLE_ROWS) * _abi.MAX_MATERIALS),
                ("photo", (C.c_double * _abi.TABLE_ROWS) * _abi.MAX_MATERIALS),
                ("total", (C.c_double * _abi.TABLE_ROWS) * _abi.MAX_MATERIALS),
                ("ff_points", C.c_int32),
                ("ff_x2", (C.c_double * _abi.FF_POINTS) * _abi.MAX_MATERIALS),
                ("ff_cum", (C.c_double * _abi.FF_POINTS) * _abi.MAX_MATERIALS)]


class OracleOpts(C.Structure):
    _fields_ = [("rng_mode", C.c_int32), ("quirks", C.c_int32), ("seed", C.c_uint64),
                ("preadvance_start", C.c_double), ("track_box", C.c_double * 3),
                ("n_threads", C.c_int32), ("label_center", C.c_double * 3),
                ("clear_grid", C.c_void_p), ("clear_dims", C.c_int32 * 3), ("heavy", C.c_int32)]


class OracleResult(C.Structure):
    _fields_ = [("histories", C.c_uint64), ("primaries", C.c_uint64), ("scatter_detected", C.c_uint64),
                ("absorbed", C.c_uint64), ("interactions", C.c_uint64), ("coherent", C.c_uint64),
                ("compton", C.c_uint64), ("woodcock_steps", C.c_uint64),
                ("num_scatter", C.c_int64), ("num_nd", C.c_uint64),
                ("sum_e_primary", C.c_double), ("sum_e_scatter", C.c_double)]


def tables_from_arrays(mats):
    """mats: list of (float64 [4][201] coh/compton/photo/total, density)."""
    t = OracleTables()
    t.n_materials = len(mats)
    for m, (a, rho) in enumerate(mats):
        t.density[m] = rho
        for k in range(_abi.TABLE_ROWS):
            t.coh[m][k], t.compt[m][k], t.photo[m][k], t.total[m][k] = a[0, k], a[1, k], a[2, k], a[3, k]
    return t


def tables_from_xs(xs):
    """The float tables of a monte_mc_xs widened to double: oracle and CUDA path then use
    identical table values."""
    t = OracleTables()
    t.n_materials = xs.n_materials
    for m in range(xs.n_materials):
        t.density[m] = xs.density[m]
        for k in range(_abi.TABLE_ROWS):
            t.coh[m][k], t.compt[m][k] = xs.coh[m][k], xs.compt[m][k]
            t.photo[m][k], t.total[m][k] = xs.photo[m][k], xs.total[m][k]
        for i in range(xs.ff_points):
            t.ff_x2[m][i], t.ff_cum[m][i] = xs.ff_x2[m][i], xs.ff_cum[m][i]
    t.ff_points = xs.ff_points
    return t


def mc_opts(rng_mode=RNG_PHILOX, quirks=0, seed=1, n_threads=0):
    o = OracleOpts()
    o.rng_mode, o.quirks, o.seed, o.n_threads = rng_mode, quirks, seed, n_threads
    o.preadvance_start = 6.0
    o.track_box[0], o.track_box[1], o.track_box[2] = 62.0, 62.0, 17.0
    o.label_center[0], o.label_center[1], o.label_center[2] = 90.0, 90.0, 160.0
    return o


def mc_run(g, vol, labels, tables, spec, opts, per, views=None, n_range=None, pixels=None, want_fates=False):
    """Oracle transport.  Returns (image0, image5 [n_views][ny][nx] int32, result dict, fates|None, fate_e|None)."""
    labels = np.ascontiguousarray(labels, np.uint8)
    vb, ve = views if views else (0, g.n_views)
    nb, ne = n_range if n_range else (0, per)
    i0, i1, j0, j1 = pixels if pixels else (0, g.ny, 0, g.nx)
    im0 = np.zeros((g.n_views, g.ny, g.nx), np.int32)
    im5 = np.zeros((g.n_views, g.ny, g.nx), np.int32)
    res = OracleResult()
    fates = fe = None
    if want_fates:
        assert ve - vb == 1
        fates = np.zeros(g.ny * g.nx * per, np.uint32)
        fe = np.zeros(g.ny * g.nx * per, np.float32)
    fn = lib().oracle_mc_run
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(vol), labels.ctypes.data_as(C.c_void_p), C.byref(tables),
            C.byref(spec) if spec is not None else None, C.byref(opts),
            C.c_uint32(per), C.c_uint32(nb), C.c_uint32(ne), C.c_int(vb), C.c_int(ve),
            C.c_int(i0), C.c_int(i1), C.c_int(j0), C.c_int(j1),
            im0.ctypes.data_as(C.c_void_p), im5.ctypes.data_as(C.c_void_p), C.byref(res),
            fates.ctypes.data_as(C.c_void_p) if want_fates else None,
            fe.ctypes.data_as(C.c_void_p) if want_fates else None)
    assert rc == 0
    return im0, im5, {k: getattr(res, k) for k, _ in res._fields_}, fates, fe


def with_clearance(opts, grid, heavy):
    """attach the clearance grid of monte_mc_clearance_grid (uint8 [gz][gy][gx]) to oracle options; the volume's
    tracking_mode switches the two-level majorant on.  Returns (opts, array to keep alive)."""
    grid = np.ascontiguousarray(grid, np.uint8)                   # [gz][gy][gx], or [8][gz][gy][gx] for DIRECTIONAL
    opts.clear_grid = grid.ctypes.data
    opts.clear_dims[0], opts.clear_dims[1], opts.clear_dims[2] = grid.shape[-1], grid.shape[-2], grid.shape[-3]
    opts.heavy = heavy
    return opts, grid


def project_primary(g, vol, labels, tables, keV, views=None):
    labels = np.ascontiguousarray(labels, np.uint8)
    vb, ve = views if views else (0, g.n_views)
    out = np.zeros((g.n_views, g.ny, g.nx), np.float32)
    lib().oracle_project_primary(C.byref(g), C.byref(vol), labels.ctypes.data_as(C.c_void_p), C.byref(tables),
                                 C.c_double(keV), C.c_int(vb), C.c_int(ve), out.ctypes.data_as(C.c_void_p))
    return out


def counts_to_map(counts, per):
    counts = np.ascontiguousarray(counts, np.int32)
    out = np.empty(counts.shape, np.float32)
    lib().oracle_counts_to_map(counts.ctypes.data_as(C.c_void_p), C.c_size_t(counts.size), C.c_int32(per),
                               out.ctypes.data_as(C.c_void_p))
    return out


def philox2x32(c0, c1, key):
    out = (C.c_uint32 * 2)()
    lib().oracle_philox2x32(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(key), out)
    return out[0], out[1]


def ref_cbct_real2(timeout=600):
    """Run the UNMODIFIED monte_cpp/CBCT_real2.cpp binary as shipped (1 pixel, 1 view, 1e7 photons,
    scatter-only tally) with its time() seed pinned to 5489 by oracle/shim/windows.h.
    Inputs: spher01.raw from the reference's own make_image01, the reference's xcom2.csv / Ca.csv
    (read in place from /root/reference), dummy 250-row spectrum CSVs (the energy is forced to 140
    at CBCT_real2.cpp:188).  Returns (image [65][65] int32, dict of the printed counters)."""
    import re
    tmp = tempfile.mkdtemp(prefix="monte_refmc_")
    try:
        subprocess.check_call([os.path.join(REF_DIR, "make_image01")], cwd=tmp)
        for fn, key in (("xcom2.csv", "h2o"), ("Ca.csv", "ca")):
            src = os.path.join(REFERENCE_ROOT, "monte_cpp", fn)
            if os.path.exists(src):
                shutil.copy(src, os.path.join(tmp, fn))
            else:
                # no reference tree (GPU box): rewrite the CSV from the packed copy of the same tables
                # (scripts/make_xs_tables.py); %.17g round-trips every double, and the reader overwrites
                # the first field of row 1 anyway (CBCT_real2.cpp:663), so the binary computes the same
                t = np.load(os.path.join(os.path.dirname(HERE), "monte_b200", "data", "xs_tables.npz"))[key]
                with open(os.path.join(tmp, fn), "w") as f:
                    f.write("\n".join("%.17g,%.17g,%.17g,%.17g" % tuple(t[:, k]) for k in range(1, 201)))
        rows = "\n".join("%g,%g,%g,%g" % (1.0, 1.0, 1.0, (i + 1) / 250.0) for i in range(250))
        for fn in ("125kv_al2mm.csv", "125kv_al10mm.csv"):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(rows)
        p = subprocess.run([os.path.join(REF_DIR, "CBCT_real2")], cwd=tmp, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=timeout)
        out = p.stdout.decode("utf-8", "replace")
        img = np.fromfile(os.path.join(tmp, "CBCTtest3\\monte005.raw"), np.int32).reshape(65, 65)
        with open(os.path.join(tmp, "spher01.raw"), "rb") as f:
            sphere = np.frombuffer(f.read(), np.uint8).reshape(325, 185, 185).copy()
        c = {"num_scatter": int(re.search(r"num_scatter (-?\d+)", out).group(1)),
             "num_nd": int(re.search(r"num not ditected (\d+)", out).group(1)),
             "count": int(re.search(r"count = (\d+)", out).group(1))}
        return img, c, sphere
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def rayleigh_round(tables, material, keV, r_a, r_acc):
    """one rejection round of the form-factor sampler: (accepted, cos_theta)"""
    c = C.c_double(0.0)
    fn = lib().oracle_rayleigh_round
    fn.restype = C.c_int
    ok = fn(C.byref(tables), C.c_int(material), C.c_double(keV), C.c_double(r_a), C.c_double(r_acc), C.byref(c))
    return bool(ok), c.value
