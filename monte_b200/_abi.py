"""ctypes mirror of include/monte_gpu.h (struct layouts and prototypes).

Python is test/bench plumbing here: the product is libmonte_gpu.so (CUDA, C ABI) and the
C++ drivers under monte_b200/host/.  Field order and types must match the header exactly;
tests/test_abi.py checks sizeof() of every struct against the library's own numbers.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmonte_gpu.so")

MAX_MATERIALS = 8
TABLE_ROWS = 201
FF_POINTS = 128
STATS_WORDS = 16

FDK_REFERENCE, FDK_TEXTBOOK = 0, 1
COORD_SCALE_AFTER, COORD_SCALE_BEFORE = 0, 1
SOURCE_PENCIL, SOURCE_CONE = 0, 1
COHERENT_FORWARD, COHERENT_FORMFACTOR = 0, 1
TRACK_GLOBAL, TRACK_CLEARANCE, TRACK_AUTO, TRACK_ADAPTIVE, TRACK_DIRECTIONAL = 0, 1, 2, 3, 4
MAJORANT_ALL, MAJORANT_PRESENT = 0, 1
DETECTOR_FLAT, DETECTOR_RING = 0, 1


class HuClass(C.Structure):
    _fields_ = [("hu_min", C.c_float), ("material_a", C.c_int32), ("material_b", C.c_int32),
                ("frac_b", C.c_float), ("density", C.c_float)]


class FdkGeom(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("nu", C.c_int32), ("nv", C.c_int32),
        ("du", C.c_double), ("dv", C.c_double),
        ("half_u", C.c_double), ("half_v", C.c_double),
        ("dso", C.c_double), ("dsd", C.c_double),
        ("weight_dist", C.c_double), ("filter_scale", C.c_double),
        ("out_scale", C.c_double), ("out_scale2", C.c_double),
        ("angle0_deg", C.c_double), ("angle_step_deg", C.c_double),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("vox", C.c_double), ("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double),
        ("s_begin", C.c_int32), ("s_end", C.c_int32),
        ("t_begin", C.c_int32), ("t_end", C.c_int32),
        ("z_begin", C.c_int32), ("z_end", C.c_int32),
        ("mask_cs", C.c_int32), ("mask_ct", C.c_int32), ("mask_cz", C.c_int32),
        ("mask_r2", C.c_int64),
        ("weight_mode", C.c_int32), ("coord_mode", C.c_int32),
    ]

    def copy(self):
        g = FdkGeom()
        C.memmove(C.byref(g), C.byref(self), C.sizeof(FdkGeom))
        return g

    def full_roi(self):
        self.s_begin, self.s_end = 0, self.nx
        self.t_begin, self.t_end = 0, self.ny
        self.z_begin, self.z_end = 0, self.nz
        return self


class FdkStats(C.Structure):
    _fields_ = [
        ("ms_h2d", C.c_double), ("ms_filter", C.c_double), ("ms_backproject", C.c_double),
        ("ms_transpose", C.c_double), ("ms_d2h", C.c_double), ("ms_total", C.c_double),
        ("voxel_updates", C.c_uint64), ("filter_macs", C.c_uint64),
        ("launches", C.c_int32), ("sm_count", C.c_int32),
    ]


class McXs(C.Structure):
    _fields_ = [
        ("n_materials", C.c_int32),
        ("density", C.c_float * MAX_MATERIALS),
        ("coh", (C.c_float * TABLE_ROWS) * MAX_MATERIALS),
        ("compt", (C.c_float * TABLE_ROWS) * MAX_MATERIALS),
        ("photo", (C.c_float * TABLE_ROWS) * MAX_MATERIALS),
        ("total", (C.c_float * TABLE_ROWS) * MAX_MATERIALS),
        ("ff_points", C.c_int32),
        ("ff_x2", (C.c_float * FF_POINTS) * MAX_MATERIALS),
        ("ff_cum", (C.c_float * FF_POINTS) * MAX_MATERIALS),
    ]


class McVolume(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("pitch", C.c_double),
        ("origin", C.c_double * 3),
        ("clip_lo", C.c_double * 3), ("clip_hi", C.c_double * 3),
        ("tracking_mode", C.c_int32), ("clearance_cell_log2", C.c_int32),
        ("majorant_mode", C.c_int32), ("reserved0", C.c_int32),
    ]


class McGeom(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32),
        ("angle0_deg", C.c_double), ("angle_step_deg", C.c_double),
        ("ny", C.c_int32), ("nx", C.c_int32),
        ("pixel", C.c_double), ("half", C.c_double),
        ("dso", C.c_double), ("dod", C.c_double),
        ("source_mode", C.c_int32), ("max_scatter", C.c_int32),
        ("detector_mode", C.c_int32), ("coherent_mode", C.c_int32),
        ("detector_shape", C.c_int32), ("reserved1", C.c_int32), ("ring_radius", C.c_double),
    ]


class McSpectrum(C.Structure):
    _fields_ = [
        ("n_bins", C.c_int32), ("bin_keV", C.c_double), ("mono_keV", C.c_double),
        ("cdf", C.POINTER(C.c_float)),
    ]


class McStats(C.Structure):
    _fields_ = [
        ("histories", C.c_uint64), ("primaries", C.c_uint64), ("scatter_detected", C.c_uint64),
        ("absorbed", C.c_uint64), ("interactions", C.c_uint64),
        ("coherent", C.c_uint64), ("compton", C.c_uint64), ("woodcock_steps", C.c_uint64),
        ("sum_e_primary", C.c_double), ("sum_e_scatter", C.c_double),
        ("ms_h2d", C.c_double), ("ms_kernel", C.c_double), ("ms_d2h", C.c_double),
        ("ms_total", C.c_double),
        ("launches", C.c_int32), ("sm_count", C.c_int32),
    ]


def stats_dict(st):
    return {k: getattr(st, k) for k, _ in st._fields_}


def bp3d20_geom():
    """Literals of recon/bp3d20.cpp (same as monte_fdk_geom_bp3d20 in the library)."""
    g = FdkGeom()
    g.n_views = 360
    g.nu = g.nv = 65
    g.du = g.dv = 0.5
    g.half_u = g.half_v = 16.25
    g.dso, g.dsd, g.weight_dist = 160.0, 220.0, 60.0
    g.filter_scale, g.out_scale, g.out_scale2 = 0.5, 2.7, 1.0
    g.angle0_deg, g.angle_step_deg = 0.0, 1.0
    g.nx = g.ny = g.nz = 256
    g.vox = 0.1
    g.x0, g.y0, g.z0 = -12.8, 12.8, 12.8
    g.s_begin, g.s_end = 125, 130
    g.t_begin, g.t_end = 0, 256
    g.z_begin, g.z_end = 0, 256
    g.mask_cs = g.mask_ct = g.mask_cz = 128
    g.mask_r2 = 118 * 118
    g.weight_mode = FDK_REFERENCE
    g.coord_mode = COORD_SCALE_AFTER
    return g


def bp3d20_325_geom():
    """Literals of recon/bp3d20_325.cpp."""
    g = bp3d20_geom()
    g.nu = g.nv = 325
    g.du = g.dv = 0.1
    g.out_scale2 = 5.0
    g.mask_cs = g.mask_ct = g.mask_cz = 0
    g.mask_r2 = -1
    g.coord_mode = COORD_SCALE_BEFORE
    return g


def fbp2_geom():
    """Literals of recon/fbp2.cpp."""
    g = bp3d20_geom()
    g.nu, g.nv = 65, 1
    g.nz = 1
    g.z_begin, g.z_end = 0, 1
    g.s_begin, g.s_end = 0, 256
    g.mask_cs = g.mask_ct = g.mask_cz = 0
    g.mask_r2 = -1
    g.dv = 0.5
    g.out_scale = 1.7
    return g


def generic_fdk_geom(n_views, nu, nv, n, *, full_circle=True, textbook=False):
    """A consistent scaled geometry for the BASELINE sweep sizes (C3/C5): the detector of
    nu x nv pixels spans the reference's 32.5 cm x (32.5*nv/nu) cm, the n^3 volume spans
    25.6 cm, views are spread over 360 degrees."""
    g = bp3d20_geom()
    g.n_views = n_views
    g.nu, g.nv = nu, nv
    g.du = 32.5 / nu
    g.dv = g.du
    g.half_u = 0.5 * nu * g.du
    g.half_v = 0.5 * nv * g.dv
    g.angle_step_deg = 360.0 / n_views if full_circle else 1.0
    g.nx = g.ny = g.nz = n
    g.vox = 25.6 / n
    g.x0, g.y0, g.z0 = -12.8, 12.8, 12.8
    g.mask_r2 = -1
    g.full_roi()
    g.weight_mode = FDK_TEXTBOOK if textbook else FDK_REFERENCE
    g.coord_mode = COORD_SCALE_AFTER
    return g


def fan_beam_2d_geom(n_views, nu, n, *, textbook=True):
    """2-D fan-beam reconstruction as the degenerate nv = 1 case of the cone-beam kernels: one detector row
    in the central plane, one slice at Z = 0.  half_v = 0 puts every voxel's axial detector coordinate exactly on row 0
    (x = (half_v - w) / dv = 0 with w = k Z = 0), so the bilinear fetch reduces to linear interpolation along u, and the
    cone-beam weights reduce to the fan-beam ones.  (recon/fbp2.cpp itself -- nearest-neighbour lookup, double
    arithmetic -- is reproduced by monte_gpu_fbp2.)"""
    g = generic_fdk_geom(n_views, nu, 1, n, textbook=textbook)
    g.nz = 1
    g.z_begin, g.z_end = 0, 1
    g.z0 = 0.0
    g.half_v = 0.0
    return g
