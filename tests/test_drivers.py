"""CPU: the C++ drivers build against the C ABI and their GPU-free outputs equal the reference
binaries' (make_fantom.cpp, make_image01.cpp) byte for byte."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from monte_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bins():
    build.build_lib()
    return {os.path.basename(p): p for p in build.build_drivers()}


def test_all_drivers_build(bins):
    assert set(bins) == {"make_fantom", "ctnum_to_mu", "cbct_mc", "cbct_fdk"}


def test_make_fantom_outputs(bins, tmp_path, oracle):
    d = str(tmp_path)
    subprocess.check_call([bins["make_fantom"], "disc", os.path.join(d, "disc.raw")])
    disc = np.fromfile(os.path.join(d, "disc.raw"), np.uint8).reshape(65, 65)
    jj, kk = np.ogrid[:65, :65]
    assert np.array_equal(disc, (((jj - 32) ** 2 + (kk - 46) ** 2) <= 100).astype(np.uint8))   # make_fantom.cpp:12
    subprocess.check_call([bins["make_fantom"], "sphere", os.path.join(d, "sph.raw")])
    sph = np.fromfile(os.path.join(d, "sph.raw"), np.uint8)
    assert sph.size == 185 * 185 * 325 and int(sph.sum()) == 523305        # golden: tests/golden/mc_real2.npz sphere_voxels
    if oracle.have_ref("make_fantom"):
        subprocess.check_call([os.path.join(oracle.REF_DIR, "make_fantom")], cwd=d)
        assert open(os.path.join(d, "ball_fan.raw"), "rb").read() == disc.tobytes()
        subprocess.check_call([os.path.join(oracle.REF_DIR, "make_image01")], cwd=d)
        assert hashlib.sha256(open(os.path.join(d, "spher01.raw"), "rb").read()).digest() == hashlib.sha256(sph.tobytes()).digest()
    subprocess.check_call([bins["make_fantom"], "cylinder", "33", "1.0", os.path.join(d, "cyl.raw")])
    from monte_b200 import scenes
    assert np.array_equal(np.fromfile(os.path.join(d, "cyl.raw"), np.uint8).reshape(33, 33, 33), scenes.cylinder_phantom(33, 1.0))


def test_gpu_drivers_fail_loudly_without_a_device(bins, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = str(tmp_path)
    np.zeros(360 * 65 * 65, np.float32).tofile(os.path.join(d, "map.raw"))
    p = subprocess.run([bins["cbct_fdk"], "bp3d20", os.path.join(d, "map.raw")], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"no CUDA device" in p.stderr


def _write_csv(path, table):
    with open(path, "wb") as f:
        f.write(b"\xef\xbb\xbf" + b"\r\n".join(b"%.10g,%.10g,%.10g,%.10g" % tuple(table[:, k]) for k in range(1, 201)) + b"\r\n")


@pytest.mark.gpu
def test_cbct_fdk_driver_reproduces_reference_golden(bins, tmp_path):
    """the recon main() replacement: reads the reference's input file layout, writes its output layout"""
    from conftest import GOLDEN
    d = str(tmp_path)
    gold = np.load(os.path.join(GOLDEN, "fdk_bp3d20.npz"))
    np.random.default_rng(int(gold["seed"])).random((360, 65, 65), dtype=np.float32).tofile(os.path.join(d, "mapg_0_20000h2oCa.raw"))
    out = subprocess.run([bins["cbct_fdk"], "bp3d20", os.path.join(d, "mapg_0_20000h2oCa.raw"), "t"], cwd=d,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    assert "filtered" in out                                   # the reference prints this after step 2 (bp3d20.cpp:75)
    xy = np.fromfile(os.path.join(d, "xy_t.raw"), np.float32).reshape(256, 256, 256)
    zy = np.fromfile(os.path.join(d, "zy_t.raw"), np.float32).reshape(256, 256, 256)
    f = np.fromfile(os.path.join(d, "map_t.raw"), np.float32).reshape(360, 65, 65)
    slab = xy[:, :, 125:130][::4, ::4, :]
    assert np.abs(slab - gold["slab_sub"]).max() <= 1e-4 * np.abs(gold["slab_sub"]).max()
    assert np.abs(f[gold["views_kept"]] - gold["filtered_views"]).max() <= 1e-4 * np.abs(gold["filtered_views"]).max()
    assert np.array_equal(zy.transpose(2, 1, 0), xy)


@pytest.mark.gpu
def test_cbct_mc_driver_writes_reference_layouts(bins, tmp_path):
    from monte_b200 import scenes
    d = str(tmp_path)
    h2o, ca = scenes.load_tables()
    _write_csv(os.path.join(d, "xcom2.csv"), h2o)
    _write_csv(os.path.join(d, "Ca.csv"), ca)
    subprocess.check_call([bins["make_fantom"], "cylinder", "33", "1.0", os.path.join(d, "cyl.raw")])
    out = subprocess.run([bins["cbct_mc"], "cyl.raw", "33", "1.0", "xcom2.csv", "Ca.csv", "17", str(32.5 / 17), "4", "200", "3", "t"],
                         cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    assert "histories" in out
    n = 4 * 17 * 17
    p0 = np.fromfile(os.path.join(d, "proj_t0.raw"), np.int32)
    p5 = np.fromfile(os.path.join(d, "proj_t5.raw"), np.int32)
    m0 = np.fromfile(os.path.join(d, "map_t0.raw"), np.float32)
    m5 = np.fromfile(os.path.join(d, "map_t5.raw"), np.float32)
    assert p0.size == p5.size == m0.size == m5.size == n
    assert (p5 >= p0).all() and p0.max() <= 200 and p0.reshape(4, 17, 17)[:, 0, :].min() == 200   # edge rays miss the phantom
    ref = -np.log(np.clip(p0, 1, 200).astype(np.float64)) + np.log(200.0)
    assert np.abs(m0 - ref).max() < 1e-5 and (m5 <= m0 + 1e-6).all()
