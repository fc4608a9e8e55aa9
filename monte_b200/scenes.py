"""Synthetic scenes for tests and bench.py: cross-section tables, phantoms, MC geometries.

Everything here is input preparation on the host (numpy); transport and reconstruction run in
libmonte_gpu.  Shapes follow BASELINE.json's configs and SURVEY.md §8(d).
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import McGeom, McSpectrum, McVolume, McXs

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "xs_tables.npz")


def load_tables(quirk_bom=False):
    """(h2o, ca) float64 [4][201] = coh, compton, photo, total; index = keV
    (monte_cpp/xcom2.csv, Ca.csv as packed by scripts/make_xs_tables.py)."""
    z = np.load(DATA)
    h2o, ca = z["h2o"].copy(), z["ca"].copy()
    if quirk_bom:                       # CBCT_real2.cpp:663, applied to every file
        h2o[0, 1] = ca[0, 1] = 1.372
    return h2o, ca


def make_xs(materials=("h2o", "ca"), quirk_bom=False):
    """monte_mc_xs with water (label 1, rho 1.0) and calcium (label 2, rho 1.55)
    (CBCT_real325im.cu:115-117)."""
    h2o, ca = load_tables(quirk_bom)
    # "pmma": the reference's third material (CBCT_real325im.cu:117, density 1.18); its PMMA.txt is not in
    # the repository, so the water table stands in for the mass coefficients
    tabs = {"h2o": (h2o, 1.0), "ca": (ca, 1.55), "pmma": (h2o, 1.18)}
    xs = McXs()
    xs.n_materials = len(materials)
    for m, name in enumerate(materials):
        t, rho = tabs[name]
        xs.density[m] = rho
        for k in range(_abi.TABLE_ROWS):
            xs.coh[m][k], xs.compt[m][k], xs.photo[m][k], xs.total[m][k] = t[0, k], t[1, k], t[2, k], t[3, k]
    return xs


# x0 (1/Angstrom) of the analytic hydrogen-like form factor F^2 ~ (1 + x^2/x0^2)^-4 used where Rayleigh deflection
# is switched on (monte_xs_formfactor_hydrogenic; the reference ships no form-factor data): 0.30 * Z_eff
FF_X0 = {"h2o": 1.0, "ca": 2.2, "pmma": 0.9}


def add_formfactors(xs, materials=("h2o", "ca")):
    """fill the Rayleigh form-factor tables of `xs` (needed for McGeom.coherent_mode = COHERENT_FORMFACTOR);
    a host helper of libmonte_gpu, no device needed"""
    from . import api
    lib = api.load()
    for m, name in enumerate(materials):
        rc = lib.monte_xs_formfactor_hydrogenic(C.byref(xs), m, FF_X0[name])
        if rc != 0:
            raise api.MonteError(lib.monte_gpu_last_error().decode())
    return xs


def cylinder_phantom(n, pitch, radius=10.0, half_len=10.0, rods=True, rod_r=1.5, rod_ring=5.0):
    """Water cylinder (axis z) with 8 calcium rods every 45 degrees — the analytic phantom of
    CBCT_real325.cu:916-921 re-oriented as in CBCT_real325im.cu:903-904, voxelised to uint8 labels
    (0 air, 1 H2O, 2 Ca) on an n^3 grid centred on the rotation axis, x fastest."""
    c = (np.arange(n) + 0.5) * pitch - 0.5 * n * pitch
    x, y, z = c[None, None, :], c[None, :, None], c[:, None, None]
    lab = np.zeros((n, n, n), np.uint8)
    inside = (x * x + y * y <= radius * radius) & (np.abs(z) <= half_len)
    lab[np.broadcast_to(inside, lab.shape)] = 1
    if rods:
        for a in range(8):
            cx, cy = rod_ring * np.cos(a * np.pi / 4), rod_ring * np.sin(a * np.pi / 4)
            rod = ((x - cx) ** 2 + (y - cy) ** 2 <= rod_r * rod_r) & (np.abs(z) <= half_len)
            lab[np.broadcast_to(rod, lab.shape)] = 2
    return lab


def hu_head_phantom(n, pitch, cortical=False):
    """A synthetic CT volume in HU, float32 [n][n][n] (z, y, x): an ellipsoidal head of soft tissue (40 HU) in air
    (-1000) with a skull shell (900 HU; cortical=True: 1400-HU plates at the sides), a fat layer under the skin (-90), two air
    sinuses, a water-filled ventricle (5) and a lung-like insert (-750).  Values sit well inside the default classes
    of monte_hu_classes_default (air | lung | adipose | soft | muscle | spongy | bone | dense | cortical)."""
    c = (n - 1) / 2.0
    z, y, x = np.meshgrid(*(((np.arange(n) - c) * pitch,) * 3), indexing="ij")
    R = 0.42 * n * pitch
    r_out = np.sqrt((x / R) ** 2 + (y / (0.85 * R)) ** 2 + (z / (0.95 * R)) ** 2)
    hu = np.full((n, n, n), -1000.0, np.float32)
    hu[r_out <= 1.0] = -90.0                     # skin + fat
    hu[r_out <= 0.93] = 900.0                    # skull
    if cortical:
        hu[(r_out <= 0.93) & (np.abs(x) > 0.7 * R)] = 1400.0
    hu[r_out <= 0.82] = 40.0                     # brain
    hu[(x / (0.25 * R)) ** 2 + (y / (0.12 * R)) ** 2 + (z / (0.2 * R)) ** 2 <= 1.0] = 5.0          # ventricle
    for sx in (-1.0, 1.0):
        hu[((x - sx * 0.3 * R) / (0.12 * R)) ** 2 + ((y - 0.45 * R) / (0.15 * R)) ** 2 + (z / (0.2 * R)) ** 2 <= 1.0] = -1000.0
    hu[(x / (0.2 * R)) ** 2 + ((y + 0.4 * R) / (0.15 * R)) ** 2 + ((z - 0.3 * R) / (0.2 * R)) ** 2 <= 1.0] = -750.0
    hu[((x / (0.2 * R)) ** 2 + ((y + 0.4 * R) / (0.15 * R)) ** 2 + ((z + 0.3 * R) / (0.2 * R)) ** 2 <= 1.0)] = 100.0   # muscle-like
    return hu


def volume_for(labels, pitch, tight=True):
    """monte_mc_volume centred on the origin; clip box = bounding box of the non-air voxels
    (outside it the photon flies straight, CBCT_real2.cpp:770)."""
    nz, ny, nx = labels.shape
    v = McVolume()
    v.nx, v.ny, v.nz = nx, ny, nz
    v.pitch = pitch
    org = (-0.5 * nx * pitch, -0.5 * ny * pitch, -0.5 * nz * pitch)
    for a in range(3):
        v.origin[a] = org[a]
    lo, hi = [0, 0, 0], [nx, ny, nz]
    if tight and labels.any():
        nzv = np.nonzero(labels)
        for a, ax in enumerate((2, 1, 0)):       # x,y,z <- array axes 2,1,0
            lo[a], hi[a] = int(nzv[ax].min()), int(nzv[ax].max()) + 1
    for a in range(3):
        v.clip_lo[a] = org[a] + lo[a] * pitch
        v.clip_hi[a] = org[a] + hi[a] * pitch
    return v


def mc_geom(n_det, pixel, n_views=360, source_mode=_abi.SOURCE_PENCIL, max_scatter=5, ny=None, nx=None):
    """Reference CBCT geometry (CBCT_real325im.cu:459): Dso 160, Dod 60, detector 32.5 cm."""
    g = McGeom()
    g.n_views = n_views
    g.angle0_deg, g.angle_step_deg = 0.0, 360.0 / n_views if n_views != 360 else 1.0
    g.ny, g.nx = ny or n_det, nx or n_det
    g.pixel = pixel
    g.half = 16.25
    g.dso, g.dod = 160.0, 60.0
    g.source_mode = source_mode
    g.max_scatter = max_scatter
    return g


def ring_geom(n_phi, n_axial, pixel, radius, n_views=1, source_mode=_abi.SOURCE_PENCIL, max_scatter=5):
    """monte_mc_geom with a ring detector (monte_cpp/circle3_2.cpp): source at the origin, n_phi angular bins
    around the z axis at `radius`, n_axial bins of height `pixel` centred on the central plane"""
    g = mc_geom(0, pixel, n_views=n_views, source_mode=source_mode, max_scatter=max_scatter, ny=n_phi, nx=n_axial)
    g.half = 0.5 * n_axial * pixel
    g.detector_shape = _abi.DETECTOR_RING
    g.ring_radius = radius
    return g


def mono_spectrum(keV=140.0):
    s = McSpectrum()
    s.n_bins, s.bin_keV, s.mono_keV = 0, 0.5, keV
    s.cdf = None
    return s


def kramers_spectrum(kvp=120.0, bin_keV=0.5, filt_cm_h2o=2.5):
    """Synthetic 120 kVp spectrum: Kramers N(E) ~ (kVp-E)/E hardened by
    exp(-mu_H2O(E)*2.5 cm) (the repo has no aluminium table), 0.5 keV bins, as a CDF.
    Returns (McSpectrum, cdf array to keep alive)."""
    h2o, _ = load_tables()
    n = int(round(kvp / bin_keV))
    e = (np.arange(n) + 1) * bin_keV
    k = np.clip((e + 0.5).astype(int), 1, 200)
    w = np.where(e < kvp, (kvp - e) / e * np.exp(-h2o[3, k] * filt_cm_h2o), 0.0)
    w[e < 10.0] = 0.0
    cdf = np.concatenate([[0.0], np.cumsum(w) / w.sum()]).astype(np.float32)
    cdf[-1] = 1.0
    s = McSpectrum()
    s.n_bins, s.bin_keV, s.mono_keV = n, bin_keV, kvp
    s.cdf = cdf.ctypes.data_as(C.POINTER(C.c_float))
    return s, cdf


def config_c1():
    """BASELINE config 1 (CPU-runnable): 65^3 labels @0.5 cm, 65x65 detector @0.5 cm."""
    lab = cylinder_phantom(65, 0.5)
    return mc_geom(65, 0.5), volume_for(lab, 0.5), lab


def config_c2():
    """BASELINE config 2: 325^3 labels @0.1 cm, 325x325 detector @0.1 cm, 360 views."""
    lab = cylinder_phantom(325, 0.1)
    return mc_geom(325, 0.1), volume_for(lab, 0.1), lab
