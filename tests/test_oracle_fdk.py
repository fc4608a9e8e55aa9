"""CPU: the FDK oracle against the reference's own outputs.

(1) bit equality with the UNMODIFIED reference binaries (oracle/_ref) where they are built,
(2) bit equality (sha256 + sub-samples) with the committed golden outputs of those binaries,
so the oracle stays pinned on boxes without /root/reference.
"""
import hashlib
import os

import numpy as np
import pytest

from monte_b200 import _abi
from conftest import GOLDEN


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand(seed, shape):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def test_oracle_bp3d20_matches_golden(oracle):
    gold = np.load(os.path.join(GOLDEN, "fdk_bp3d20.npz"))
    g = _abi.bp3d20_geom()
    f, xy, zy = oracle.fdk(g, rand(int(gold["seed"]), (360, 65, 65)), want_zy=True)
    assert sha(f) == str(gold["filtered_sha"])
    slab = xy[:, :, 125:130]
    assert sha(slab) == str(gold["slab_sha"])
    assert np.array_equal(slab[::4, ::4, :], gold["slab_sub"])
    assert np.array_equal(f[gold["views_kept"]], gold["filtered_views"])
    assert np.array_equal(zy.transpose(2, 1, 0), xy)          # bp3d20.cpp:160-161
    assert not xy[:, :, :125].any() and not xy[:, :, 130:].any()   # shipped loop: s in [125,130)


def test_oracle_fbp2_matches_golden(oracle):
    gold = np.load(os.path.join(GOLDEN, "fdk_fbp2.npz"))
    f, img = oracle.fbp2(_abi.fbp2_geom(), rand(int(gold["seed"]), (360, 65)), view_first=1)
    assert sha(f) == str(gold["filtered_sha"])
    assert sha(img) == str(gold["image_sha"])


@pytest.mark.slow
def test_oracle_bp3d20_325_matches_golden(oracle):
    gold = np.load(os.path.join(GOLDEN, "fdk_bp3d20_325.npz"))
    g = _abi.bp3d20_325_geom()
    f, xy, _ = oracle.fdk(g, rand(int(gold["seed"]), (360, 325, 325)))
    assert sha(f) == str(gold["filtered_sha"])
    slab = xy[:, :, 125:130].copy()
    u = gold["undefined"].astype(int)      # voxels where the binary reads past its heap buffer (Q9)
    slab[u[:, 0], u[:, 1], u[:, 2]] = 0
    assert sha(slab) == str(gold["slab_sha"])


def test_oracle_equals_unmodified_reference_binary(oracle):
    if not oracle.have_ref("bp3d20"):
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    proj = rand(11, (360, 65, 65))
    f_ref, xy_ref, zy_ref = oracle.ref_bp3d20(proj)
    f, xy, zy = oracle.fdk(_abi.bp3d20_geom(), proj, want_zy=True)
    assert np.array_equal(f, f_ref)
    assert np.array_equal(xy, xy_ref)
    assert np.array_equal(zy, zy_ref)


def test_oracle_fbp2_equals_unmodified_reference_binary(oracle):
    if not oracle.have_ref("fbp2"):
        pytest.skip("oracle/_ref not built")
    sino = rand(12, (360, 65))
    f_ref, img_ref = oracle.ref_fbp2(sino)
    f, img = oracle.fbp2(_abi.fbp2_geom(), sino, 1)
    assert np.array_equal(f, f_ref) and np.array_equal(img, img_ref)


def test_ramp_taps(oracle):
    """bp3d20.cpp:48-60: h[0]=1/4, odd n -> -1/(n pi)^2, even -> 0."""
    import ctypes as C
    nu = 65
    r = np.zeros(2 * nu - 1, np.float32)
    oracle.lib().oracle_fdk_ramp(nu, r.ctypes.data_as(C.POINTER(C.c_float)))
    assert r[nu - 1] == np.float32(0.25)
    assert r[nu] == np.float32(-1.0 / np.pi ** 2) and r[nu - 2] == r[nu]
    assert r[nu + 1] == 0 and r[nu + 2] == np.float32(-1.0 / (3 * np.pi) ** 2)


def test_filter_linearity_and_empty_roi(oracle):
    g = _abi.bp3d20_geom()
    g.n_views = 3
    a, b = rand(1, (3, 65, 65)), rand(2, (3, 65, 65))
    fa, fb, fab = oracle.fdk_filter(g, a), oracle.fdk_filter(g, b), oracle.fdk_filter(g, a + b)
    assert np.allclose(fa + fb, fab, atol=2e-6)
    g.s_begin = g.s_end = 0                      # empty region: nothing is reconstructed
    _, vol, _ = oracle.fdk(g, a)
    assert not vol.any()
