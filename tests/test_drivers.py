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
