"""Thin ctypes layer over libmonte_gpu.so for tests, bench.py and the multi-GPU plumbing.

There is no CPU fallback: if the library cannot be loaded, or no B200 is visible at init time,
every entry point raises MonteError.  The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (FdkGeom, FdkStats, McGeom, McSpectrum, McStats, McVolume, McXs)


class MonteError(RuntimeError):
    pass


_lib = None


def load():
    """Load libmonte_gpu.so (no device needed) and declare the prototypes of include/monte_gpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_abi.LIB_PATH):
        raise MonteError("libmonte_gpu.so not built (%s); run `python -m monte_b200.build` — "
                         "there is no CPU fallback" % _abi.LIB_PATH)
    _lib = _bind(_abi.LIB_PATH)
    return _lib


def _bind(path):
    """Open the shared library at `path` and declare the prototypes of include/monte_gpu.h on it.  load() binds the CUDA
    library; the only other caller is tests/emu/build.py (CPU emulation of the same sources, tests only)."""
    lib = C.CDLL(path)
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    G = C.POINTER(FdkGeom)
    proto = {
        "monte_gpu_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
        "monte_gpu_shutdown": (None, []),
        "monte_gpu_last_error": (C.c_char_p, []),
        "monte_gpu_abi_version": (C.c_int, []),
        "monte_gpu_sm_count": (C.c_int, []),
        "monte_gpu_device_count": (C.c_int, []),
        "monte_gpu_peer_access": (C.c_int, []),
        "monte_gpu_fdk_partition": (C.c_int, [G, C.c_int, C.POINTER(C.c_int)]),
        "monte_gpu_simulate_maps": (C.c_int, [C.POINTER(McGeom), C.POINTER(McVolume), vp, C.POINTER(McXs),
                                              C.POINTER(McSpectrum), u32, u32, u32, u64, C.c_int, C.c_int, vp, vp, vp, vp,
                                              C.POINTER(McStats)]),
        "monte_fdk_geom_bp3d20": (None, [G]),
        "monte_fdk_geom_bp3d20_325": (None, [G]),
        "monte_fdk_geom_fbp2": (None, [G]),
        "monte_gpu_fdk": (C.c_int, [G, vp, vp, vp, vp, C.POINTER(FdkStats)]),
        "monte_gpu_fdk_filtered_pitch": (sz, [G]),
        "monte_gpu_fdk_filtered_elems": (sz, [G]),
        "monte_gpu_fdk_filter_dev": (C.c_int, [G, vp, C.c_int, C.c_int, vp, vp]),
        "monte_gpu_fdk_pad_dev": (C.c_int, [G, vp, vp]),
        "monte_gpu_fdk_backproject_dev": (C.c_int, [G, vp, C.c_int, C.c_int, vp, vp]),
        "monte_gpu_fdk_transpose_dev": (C.c_int, [G, vp, vp, vp]),
        "monte_gpu_fdk_backproject_views_dev": (C.c_int, [G, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
        "monte_gpu_fdk_pad_views_dev": (C.c_int, [G, vp, C.c_int, C.c_int, vp]),
        "monte_gpu_fdk_unpad_dev": (C.c_int, [G, vp, vp, vp]),
        "monte_gpu_fbp2": (C.c_int, [G, C.c_int, vp, vp, vp, C.POINTER(FdkStats)]),
        "monte_gpu_simulate": (C.c_int, [C.POINTER(McGeom), C.POINTER(McVolume), vp, C.POINTER(McXs),
                                         C.POINTER(McSpectrum), u32, u64, C.c_int, C.c_int, vp, vp,
                                         C.POINTER(McStats)]),
        "monte_gpu_simulate_range": (C.c_int, [C.POINTER(McGeom), C.POINTER(McVolume), vp, C.POINTER(McXs),
                                               C.POINTER(McSpectrum), u32, u32, u32, u64, C.c_int, C.c_int, vp, vp,
                                               C.POINTER(McStats)]),
        "monte_gpu_scene_create": (C.c_int, [C.POINTER(McGeom), C.POINTER(McVolume), vp, C.POINTER(McXs),
                                             C.POINTER(McSpectrum), C.POINTER(vp)]),
        "monte_gpu_scene_destroy": (None, [vp]),
        "monte_gpu_scene_update_labels": (C.c_int, [vp, vp, vp]),
        "monte_gpu_simulate_dev": (C.c_int, [vp, u64, C.c_int, C.c_int, u32, u32, u32, vp, vp, vp, vp]),
        "monte_gpu_mc_stats_unpack": (None, [vp, C.POINTER(McStats)]),
        "monte_gpu_simulate_fates": (C.c_int, [vp, u64, C.c_int, u32, vp, vp]),
        "monte_gpu_counts_to_map": (C.c_int, [vp, sz, i32, vp]),
        "monte_gpu_counts_to_map_dev": (C.c_int, [vp, sz, i32, vp, vp]),
        "monte_gpu_project_primary": (C.c_int, [C.POINTER(McGeom), C.POINTER(McVolume), vp, C.POINTER(McXs),
                                                C.c_double, C.c_int, C.c_int, vp]),
        "monte_gpu_projector_create": (C.c_int, [C.POINTER(McVolume), vp, C.POINTER(vp)]),
        "monte_gpu_projector_destroy": (None, [vp]),
        "monte_gpu_project_primary_dev": (C.c_int, [vp, C.POINTER(McGeom), C.POINTER(McXs), C.c_double, C.c_int, C.c_int, vp, vp]),
        "monte_xs_load_csv": (C.c_int, [C.c_char_p, C.c_int, C.c_float, C.c_int, C.POINTER(McXs)]),
        "monte_make_fantom": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "monte_make_sphere": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "monte_ctnum_to_mu": (C.c_int, [vp, sz, C.POINTER(McXs), C.c_double, C.c_float, C.c_float, vp, vp]),
        "monte_xs_majorant": (C.c_int, [C.POINTER(McXs), vp, sz, vp]),
        "monte_hu_classes_default": (C.c_int, [C.c_int, vp]),
        "monte_ctnum_segment": (C.c_int, [vp, sz, vp, C.c_int, C.POINTER(McXs), C.c_double, C.POINTER(McXs), vp, vp,
                                          C.POINTER(C.c_uint32)]),
        "monte_xs_formfactor_hydrogenic": (C.c_int, [C.POINTER(McXs), C.c_int, C.c_double]),
        "monte_mc_clearance_dims": (C.c_int, [C.POINTER(McVolume), C.c_int, C.POINTER(C.c_int32)]),
        "monte_mc_clearance_grid": (C.c_int, [C.POINTER(McVolume), vp, C.c_int, C.c_int, C.c_int, vp]),
        "monte_xs_heavy_material": (C.c_int, [C.POINTER(McXs)]),
        "monte_mc_clearance_grid_octants": (C.c_int, [C.POINTER(McVolume), vp, C.c_int, C.c_int, C.c_int, vp]),
        "monte_mc_resolve_tracking": (C.c_int, [C.POINTER(McXs), C.POINTER(McSpectrum), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
        "monte_gpu_fdk_slab_rows": (C.c_int, [C.POINTER(FdkGeom), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "monte_gpu_fdk_backproject_peers_dev": (C.c_int, [G, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp]),
        "monte_gpu_ipc_export": (C.c_int, [vp, vp, C.POINTER(C.c_uint64)]),
        "monte_gpu_ipc_open": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
        "monte_gpu_ipc_close": (C.c_int, [vp]),
    }
    missing = []
    for name, (res, args) in proto.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:           # header/library mismatch: tests/test_abi.py fails on this
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    lib._monte_symbols = sorted(proto)
    lib._monte_missing = missing
    return lib


def _check(rc):
    if rc != 0:
        raise MonteError("libmonte_gpu error %d: %s" % (rc, load().monte_gpu_last_error().decode()))


_inited_dev = None


def init(device=0):
    """Bind this process to one B200 (one process per GPU), or -- device = a list of CUDA ordinals -- to several
    GPUs of one box: the host-buffer calls simulate() / fdk() then shard over them inside the library."""
    global _inited_dev
    lib = load()
    devs = list(device) if isinstance(device, (list, tuple)) else [device]
    ids = (C.c_int * len(devs))(*devs)
    _check(lib.monte_gpu_init(len(devs), ids))
    _inited_dev = devs[0]
    return lib


def fdk_partition(g, n_parts):
    """z-slab cuts [0, ..., nz] of equal modelled work that the multi-device monte_gpu_fdk uses"""
    cuts = (C.c_int * (n_parts + 1))()
    _check(load().monte_gpu_fdk_partition(C.byref(g), n_parts, cuts))
    return list(cuts)


def shutdown():
    global _inited_dev
    if _lib is not None:
        _lib.monte_gpu_shutdown()
    _inited_dev = None


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


# ------------------------------------------------------------------ FDK, host buffers
def fdk(g, proj, want_filtered=True, want_zy=False, out=None):
    """monte_gpu_fdk on numpy (or pinned torch .numpy()) buffers.
    Returns (filtered|None, vol_xy, vol_zy|None, stats dict)."""
    lib = load()
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    if proj.shape != (g.n_views, g.nu, g.nv):
        raise MonteError("proj shape %r != (%d,%d,%d)" % (proj.shape, g.n_views, g.nu, g.nv))
    filt = np.empty((g.n_views, g.nv, g.nu), np.float32) if want_filtered else None
    vol = out if out is not None else np.empty((g.nz, g.ny, g.nx), np.float32)
    vzy = np.empty((g.nx, g.ny, g.nz), np.float32) if want_zy else None
    st = FdkStats()
    _check(lib.monte_gpu_fdk(C.byref(g), _ptr(proj), _ptr(filt), _ptr(vol), _ptr(vzy), C.byref(st)))
    return filt, vol, vzy, _abi.stats_dict(st)


def fbp2(g, sino, view_first=1):
    lib = load()
    sino = np.ascontiguousarray(sino, dtype=np.float32)
    filt = np.empty((g.n_views, g.nu), np.float32)
    img = np.empty((g.ny, g.nx), np.float32)
    st = FdkStats()
    _check(lib.monte_gpu_fbp2(C.byref(g), view_first, _ptr(sino), _ptr(filt), _ptr(img), C.byref(st)))
    return filt, img, _abi.stats_dict(st)


# ------------------------------------------------------------------ FDK, device buffers (torch)
def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def fdk_filtered_shape(g):
    lib = load()
    pitch = lib.monte_gpu_fdk_filtered_pitch(C.byref(g))
    return (g.n_views * g.nv + 2, pitch)


def fdk_filter_dev(g, d_map, d_filt, view_begin=0, view_end=None, stream=None, pad=True):
    """d_map: cuda float32 [views][nu][nv]; d_filt: cuda float32 of fdk_filtered_shape(g)."""
    lib = load()
    ve = g.n_views if view_end is None else view_end
    _check(lib.monte_gpu_fdk_filter_dev(C.byref(g), C.c_void_p(d_map.data_ptr()), view_begin, ve,
                                        C.c_void_p(d_filt.data_ptr()), _stream_ptr(stream)))
    if pad:
        _check(lib.monte_gpu_fdk_pad_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()), _stream_ptr(stream)))


def fdk_pad_dev(g, d_filt, stream=None):
    _check(load().monte_gpu_fdk_pad_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()), _stream_ptr(stream)))


def fdk_backproject_dev(g, d_filt, d_slab, z_lo=0, z_hi=None, stream=None):
    lib = load()
    zh = g.nz if z_hi is None else z_hi
    _check(lib.monte_gpu_fdk_backproject_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()), z_lo, zh,
                                             C.c_void_p(d_slab.data_ptr()), _stream_ptr(stream)))


def fdk_backproject_views_dev(g, d_filt, d_slab, z_lo, z_hi, view_lo, view_hi, continue_sum, stream=None):
    _check(load().monte_gpu_fdk_backproject_views_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()), z_lo, z_hi,
                                                      C.c_void_p(d_slab.data_ptr()), view_lo, view_hi,
                                                      1 if continue_sum else 0, _stream_ptr(stream)))


def ipc_export(d_tensor):
    """(64-byte handle, offset) of the device allocation holding a cuda tensor (or emulated device buffer), for a peer process"""
    h = (C.c_ubyte * 64)()
    off = C.c_uint64(0)
    _check(load().monte_gpu_ipc_export(C.c_void_p(d_tensor.data_ptr()), C.cast(h, C.c_void_p), C.byref(off)))
    return bytes(h), off.value


def ipc_open(handle, offset):
    """address (int) in this process of the byte a peer exported with ipc_export"""
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _check(load().monte_gpu_ipc_open(C.cast(h, C.c_void_p), offset, C.byref(p)))
    return p.value


def ipc_close(ptr):
    _check(load().monte_gpu_ipc_close(C.c_void_p(ptr)))


def fdk_backproject_peers_dev(g, seg_ptrs, seg_v_end, d_slab, z_lo, z_hi, stream=None):
    """backproject all views into d_slab = slices [z_lo, z_hi), gathering the detector-row band out of the segments'
    padded-row buffers (addresses: own tensors' data_ptr() or ipc_open() results); monte_gpu_fdk_backproject_peers_dev"""
    n = len(seg_ptrs)
    bases = (C.c_void_p * n)(*seg_ptrs)
    ends = (C.c_int * n)(*seg_v_end)
    _check(load().monte_gpu_fdk_backproject_peers_dev(C.byref(g), n, C.cast(bases, C.c_void_p), C.cast(ends, C.c_void_p), z_lo, z_hi,
                                                      C.c_void_p(d_slab.data_ptr()), _stream_ptr(stream)))


def fdk_slab_rows(g, z_lo, z_hi):
    """(row_lo, row_hi): the axial detector rows of every view that slices [z_lo, z_hi) read (beside rows 0..3)"""
    a, b = C.c_int(0), C.c_int(0)
    _check(load().monte_gpu_fdk_slab_rows(C.byref(g), z_lo, z_hi, C.byref(a), C.byref(b)))
    return a.value, b.value


def fdk_pad_views_dev(g, d_filt, view_lo, view_hi, stream=None):
    _check(load().monte_gpu_fdk_pad_views_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()), view_lo, view_hi, _stream_ptr(stream)))


def fdk_unpad_dev(g, d_filt, d_dense, stream=None):
    _check(load().monte_gpu_fdk_unpad_dev(C.byref(g), C.c_void_p(d_filt.data_ptr()),
                                          C.c_void_p(d_dense.data_ptr()), _stream_ptr(stream)))


def fdk_transpose_dev(g, d_xy, d_zy, stream=None):
    _check(load().monte_gpu_fdk_transpose_dev(C.byref(g), C.c_void_p(d_xy.data_ptr()),
                                              C.c_void_p(d_zy.data_ptr()), _stream_ptr(stream)))


# ------------------------------------------------------------------ Monte Carlo
def simulate(g, vol, labels, xs, spec, per, seed=1, views=None, n_range=None, out=None, maps=None):
    """monte_gpu_simulate(_range) on host buffers.  Returns (image0, image5 [n_views][ny][nx] int32,
    stats dict).  out=(image0, image5) reuses caller (e.g. pinned) buffers; only `views` are written.
    maps=(map0, map5) float32 arrays (or True to allocate them): monte_gpu_simulate_maps, the -log maps come
    back too, as 4th and 5th element."""
    lib = load()
    labels = np.ascontiguousarray(labels, np.uint8)
    vb, ve = views if views else (0, g.n_views)
    nb, ne = n_range if n_range else (0, per)
    if out is None:
        im0 = np.zeros((g.n_views, g.ny, g.nx), np.int32)
        im5 = np.zeros((g.n_views, g.ny, g.nx), np.int32)
    else:
        im0, im5 = out
    st = McStats()
    if maps is not None:
        m0, m5 = (np.zeros(im0.shape, np.float32), np.zeros(im0.shape, np.float32)) if maps is True else maps
        _check(lib.monte_gpu_simulate_maps(C.byref(g), C.byref(vol), _ptr(labels), C.byref(xs),
                                           C.byref(spec) if spec is not None else None, per, nb, ne, seed, vb, ve,
                                           _ptr(im0), _ptr(im5), _ptr(m0), _ptr(m5), C.byref(st)))
        return im0, im5, _abi.stats_dict(st), m0, m5
    _check(lib.monte_gpu_simulate_range(C.byref(g), C.byref(vol), _ptr(labels), C.byref(xs),
                                        C.byref(spec) if spec is not None else None, per, nb, ne, seed, vb, ve,
                                        _ptr(im0), _ptr(im5), C.byref(st)))
    return im0, im5, _abi.stats_dict(st)


class Scene:
    """Device-resident scene (labels, tables, spectrum) for repeated transport launches."""

    def __init__(self, g, vol, labels, xs, spec):
        lib = load()
        self.g = g
        self._keep = (vol, np.ascontiguousarray(labels, np.uint8), xs, spec)
        self.handle = C.c_void_p()
        _check(lib.monte_gpu_scene_create(C.byref(g), C.byref(vol), _ptr(self._keep[1]), C.byref(xs),
                                          C.byref(spec) if spec is not None else None, C.byref(self.handle)))

    def update_labels(self, labels, stream=None):
        """H2D of a new label volume (same shape) into the resident scene"""
        labels = np.ascontiguousarray(labels, np.uint8)
        _check(load().monte_gpu_scene_update_labels(self.handle, _ptr(labels), _stream_ptr(stream)))

    def close(self):
        if self.handle:
            load().monte_gpu_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def simulate_dev(self, d_image0, d_image5, per, seed=1, views=None, n_range=None, d_stats=None, stream=None):
        """Add photons n in n_range of every pixel of `views` into the cuda int32 tensors
        d_image0/d_image5 [n_views][ny][nx]; d_stats: cuda int64[16] accumulators or None."""
        vb, ve = views if views else (0, self.g.n_views)
        nb, ne = n_range if n_range else (0, per)
        _check(load().monte_gpu_simulate_dev(self.handle, seed, vb, ve, nb, ne, per,
                                             C.c_void_p(d_image0.data_ptr()), C.c_void_p(d_image5.data_ptr()),
                                             C.c_void_p(d_stats.data_ptr()) if d_stats is not None else None,
                                             _stream_ptr(stream)))

    def fates(self, view, per, seed=1):
        """Per-history fate records of one view (see include/monte_gpu.h)."""
        n = self.g.ny * self.g.nx * per
        f = np.zeros(n, np.uint32)
        e = np.zeros(n, np.float32)
        _check(load().monte_gpu_simulate_fates(self.handle, seed, view, per, _ptr(f), _ptr(e)))
        return f, e


def unpack_stats(words):
    """words: 16 uint64 (numpy) from the device accumulators -> dict"""
    st = McStats()
    w = np.ascontiguousarray(words, np.uint64)
    load().monte_gpu_mc_stats_unpack(_ptr(w), C.byref(st))
    return _abi.stats_dict(st)


def counts_to_map(counts, per):
    counts = np.ascontiguousarray(counts, np.int32)
    out = np.empty(counts.shape, np.float32)
    _check(load().monte_gpu_counts_to_map(_ptr(counts), counts.size, per, _ptr(out)))
    return out


def clearance_grid(vol, labels, xs, cell_log2=None):
    """(grid uint8 [gz][gy][gx], heavy material index) of MONTE_MC_TRACK_CLEARANCE -- host helper, no device needed"""
    lib = load()
    labels = np.ascontiguousarray(labels, np.uint8)
    cl = vol.clearance_cell_log2 if cell_log2 is None else cell_log2
    dims = (C.c_int32 * 3)()
    _check(lib.monte_mc_clearance_dims(C.byref(vol), cl, dims))
    heavy = lib.monte_xs_heavy_material(C.byref(xs))
    if vol.tracking_mode == _abi.TRACK_DIRECTIONAL:          # eight grids, one per octant of the flight direction
        grid = np.full((8, dims[2], dims[1], dims[0]), 127, np.uint8)
        if heavy >= 0:
            _check(lib.monte_mc_clearance_grid_octants(C.byref(vol), _ptr(labels), xs.n_materials, heavy, cl, _ptr(grid)))
        return grid, heavy
    grid = np.full((dims[2], dims[1], dims[0]), 127, np.uint8)
    if heavy >= 0:
        _check(lib.monte_mc_clearance_grid(C.byref(vol), _ptr(labels), xs.n_materials, heavy, cl, _ptr(grid)))
    return grid, heavy


class Projector:
    """Device-resident deterministic projector (label copies built once); project() writes line integrals into a
    device tensor [n_views][ny][nx] -- the input layout of fdk_filter_dev."""

    def __init__(self, vol, labels):
        self._labels = np.ascontiguousarray(labels, np.uint8)
        self.handle = C.c_void_p()
        _check(load().monte_gpu_projector_create(C.byref(vol), _ptr(self._labels), C.byref(self.handle)))

    def project(self, g, xs, keV, d_map, views=None, stream=None):
        vb, ve = views if views else (0, g.n_views)
        _check(load().monte_gpu_project_primary_dev(self.handle, C.byref(g), C.byref(xs), keV, vb, ve,
                                                    C.c_void_p(d_map.data_ptr()), _stream_ptr(stream)))

    def close(self):
        if self.handle:
            load().monte_gpu_projector_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hu_classes_default(have_calcium=True):
    """the library's default HU class table (monte_hu_classes_default): a ctypes array of HuClass"""
    from ._abi import HuClass
    arr = (HuClass * (_abi.MAX_MATERIALS + 1))()
    n = load().monte_hu_classes_default(1 if have_calcium else 0, C.cast(arr, C.c_void_p))
    if n < 0:
        _check(n)
    return (HuClass * n)(*arr[:n])


def ctnum_segment(hu, classes, base_xs, keV=140.0, want_mu=True):
    """N-class segmentation of a CT volume (monte_ctnum_segment, host only): returns (labels uint8 like hu, the
    McXs those labels index, mu float32 or None, present-material bit mask)"""
    hu = np.ascontiguousarray(hu, np.float32)
    labels = np.empty(hu.shape, np.uint8)
    mu = np.empty(hu.shape, np.float32) if want_mu else None
    out = McXs()
    present = C.c_uint32(0)
    _check(load().monte_ctnum_segment(_ptr(hu), hu.size, C.cast(classes, C.c_void_p), len(classes), C.byref(base_xs),
                                      float(keV), C.byref(out), _ptr(labels), _ptr(mu) if want_mu else None,
                                      C.byref(present)))
    return labels, out, mu, present.value


def resolve_tracking(xs, spec):
    """what MONTE_MC_TRACK_AUTO picks for these tables and this spectrum: (mode, cell_log2, mean majorant ratio)"""
    cl, ratio = C.c_int32(0), C.c_double(0.0)
    mode = load().monte_mc_resolve_tracking(C.byref(xs), C.byref(spec) if spec is not None else None, C.byref(cl), C.byref(ratio))
    return mode, cl.value, ratio.value


def project_primary(g, vol, labels, xs, keV, views=None, out=None):
    labels = np.ascontiguousarray(labels, np.uint8)
    vb, ve = views if views else (0, g.n_views)
    if out is None:
        out = np.zeros((g.n_views, g.ny, g.nx), np.float32)
    _check(load().monte_gpu_project_primary(C.byref(g), C.byref(vol), _ptr(labels), C.byref(xs), keV, vb, ve, _ptr(out)))
    return out
