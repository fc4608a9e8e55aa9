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
    # the CUDA-IPC form needs one address space per box: emulated ranks are separate host processes, so that block is cut
    # (its kernel path runs in-process in tests/test_emu_fdk.py::test_emu_backproject_from_segment_buffers_equals_one_buffer)
    a, b = w.index("# >>> peers"), w.index("# <<< peers")
    w = w[:a] + w[b:]
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


PIPE_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
''' + EMU_API + r'''
from monte_b200 import dist as mdist
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
api.init(0)
out = sys.argv[1]
# BASELINE config 3 in small: deterministic projection of the phantom, sharded by views (each rank projects exactly the
# views it filters), one band-limited all_to_all of filtered rows, z-slabs of equal work
lab = scenes.cylinder_phantom(33, 0.8)
vol = scenes.volume_for(lab, 0.8)
xs = scenes.make_xs()
n_views, nu, nv = 30, 40, 24
g = scenes.mc_geom(0, 32.5 / nu, n_views=n_views, ny=nu, nx=nv)
fg = _abi.generic_fdk_geom(n_views, nu, nv, 40, textbook=True)
d_map = torch.full((n_views, nu, nv), float("nan"))
filt = torch.full(api.fdk_filtered_shape(fg), float("nan"))
zr = mdist.fdk_z_partition(fg, ws, block=8)
norm = [[tuple(q)] if not isinstance(q[0], (tuple, list)) else [tuple(x) for x in q] for q in zr]
slabs = {}
pr = api.Projector(vol, lab)

def project_and_filter(a, b):
    pr.project(g, xs, 70.0, d_map, views=(a, b))
    api.fdk_filter_dev(fg, d_map, filt, a, b, pad=False)

def backproject(a, b):
    slabs[(a, b)] = torch.empty((b - a, fg.ny, fg.nx))
    api.fdk_backproject_dev(fg, filt, slabs[(a, b)], a, b)

mdist.fdk_sharded_band(project_and_filter, lambda: api.fdk_pad_dev(fg, filt), backproject, lambda a, b: api.fdk_slab_rows(fg, a, b),
                       filt, fg.n_views, fg.nv, zr)
pr.close()
np.savez(os.path.join(out, "p%%d.npz" %% rank), **{"%%d_%%d" %% k: v.numpy() for k, v in slabs.items()})
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("ws", [2, 3])
def test_emu_config3_pipeline_sharded_equals_one_process(monte_emu, tmp_path, ws):
    """projection -> filter -> band exchange -> backprojection, every stage on the "device", sharded over ws ranks:
    the slabs tile the volume and equal the single-process host-buffer chain bit for bit"""
    script = os.path.join(str(tmp_path), "pw.py")
    with open(script, "w") as f:
        f.write(PIPE_WORKER % ROOT)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % ws,
                           "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90 + ws), script, str(tmp_path)],
                          timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    m = monte_emu
    lab = scenes.cylinder_phantom(33, 0.8)
    vol = scenes.volume_for(lab, 0.8)
    xs = scenes.make_xs()
    n_views, nu, nv = 30, 40, 24
    g = scenes.mc_geom(0, 32.5 / nu, n_views=n_views, ny=nu, nx=nv)
    fg = _abi.generic_fdk_geom(n_views, nu, nv, 40, textbook=True)
    _, ref, _, _ = m.fdk(fg, m.project_primary(g, vol, lab, xs, 70.0), want_filtered=False)
    assert np.abs(ref).max() > 0
    covered = np.zeros(fg.nz, int)
    for r in range(ws):
        z = np.load(os.path.join(str(tmp_path), "p%d.npz" % r))
        for key in z.files:
            a, b = (int(x) for x in key.split("_"))
            assert np.array_equal(z[key], ref[a:b]), (r, a, b)
            covered[a:b] += 1
    assert (covered == 1).all()
