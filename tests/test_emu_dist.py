"""CPU counterpart of tests/test_dist_gpu.py: the SAME worker script (one process per "GPU", torchrun), with gloo
instead of NCCL and the emulated library (tests/emu) instead of libmonte_gpu.so.  So the multi-GPU plumbing of
monte_b200/dist.py -- photon-range split + reduce, gather / pipelined-broadcast / band-limited all_to_all
exchanges of filtered rows, z-slabs of equal work -- runs around the real kernels and C-ABI host code here, and
must reproduce the single-process result bit for bit.  Test infrastructure only (see tests/emu/cuda_runtime.h).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import test_dist_gpu as G
from monte_b200 import _abi, scenes

ROOT = G.ROOT

EMU_API = '''from monte_b200 import _abi, scenes
import importlib.util as _ilu
_spec = _ilu.spec_from_file_location("monte_emu_build", os.path.join(%r, "tests", "emu", "build.py"))
_eb = _ilu.module_from_spec(_spec); _spec.loader.exec_module(_eb)
api = _eb.api()
''' % ROOT


def emu_worker_source(ws):
    w = G.WORKER % ROOT
    if ws > 2:      # a middle slab of this small geometry reads every detector row: only the end ranks can check the band
        old = 'assert bool(torch.isnan(filt[: fg.n_views * fg.nv]).any())'
        assert old in w
        w = w.replace(old, 'assert 0 < rank < ws - 1 or bool(torch.isnan(filt[: fg.n_views * fg.nv]).any())')
    for old, new in (
            ("torch.cuda.set_device(local)\n", ""),
            ('dev = torch.device("cuda", local)', 'dev = torch.device("cpu")'),
            ('dist.init_process_group("nccl", device_id=dev)', 'dist.init_process_group("gloo")'),
            ("from monte_b200 import _abi, api, scenes\n", EMU_API),
            ("api.init(local)", "api.init(0)"),              # every emulated rank has its own one-device "box"
            ("torch.cuda.synchronize()", "pass")):
        assert old in w, old
        w = w.replace(old, new)
    assert "cuda" not in w and "nccl" not in w
    return w


@pytest.mark.parametrize("ws", [2, 3])
def test_emu_ranks_equal_one_process(monte_emu, tmp_path, ws):
    script = os.path.join(str(tmp_path), "w.py")
    with open(script, "w") as f:
        f.write(emu_worker_source(ws))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % ws,
                           "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200 + ws), script, str(tmp_path)],
                          timeout=600, env=env)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(ws)]
    lab = scenes.cylinder_phantom(33, 1.0)
    g = scenes.mc_geom(17, 32.5 / 17, n_views=3)
    g.angle_step_deg = 120.0
    r0, r5, _ = monte_emu.simulate(g, scenes.volume_for(lab, 1.0), lab, scenes.make_xs(), scenes.mono_spectrum(), 101, seed=9)
    assert np.array_equal(parts[0]["im0"], r0) and np.array_equal(parts[0]["im5"], r5)
    fg = _abi.generic_fdk_geom(45, 65, 33, 40)
    proj = np.random.default_rng(5).random((45, 65, 33), dtype=np.float32)
    _, vol, _, _ = monte_emu.fdk(fg, proj, want_filtered=False)
    z = 0
    for p in parts:
        assert p["z"][0] == z
        assert np.array_equal(p["slab"], vol[p["z"][0]:p["z"][1]])
        z = p["z"][1]
    assert z == fg.nz
    for r in range(ws):                           # the equal-work partition reconstructs the same voxels
        u = np.load(os.path.join(str(tmp_path), "u%d.npz" % r))
        assert np.array_equal(u["slab"], vol[u["z"][0]:u["z"][1]])
