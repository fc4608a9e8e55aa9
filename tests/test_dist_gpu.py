"""GPU, world_size 2, NCCL (skipped on a single-GPU box): sharded MC and FDK equal the single-GPU result."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from monte_b200 import _abi, api, scenes
from monte_b200 import dist as mdist
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
api.init(local)
out = sys.argv[1]
# MC
lab = scenes.cylinder_phantom(33, 1.0)
g = scenes.mc_geom(17, 32.5 / 17, n_views=3); g.angle_step_deg = 120.0
vol = scenes.volume_for(lab, 1.0)
sc = api.Scene(g, vol, lab, scenes.make_xs(), scenes.mono_spectrum())
im0 = torch.zeros((3, 17, 17), dtype=torch.int32, device=dev); im5 = torch.zeros_like(im0)
run = lambda a0, a5, per, views, nr: sc.simulate_dev(a0, a5, per, seed=9, views=views, n_range=nr)
for v in range(3):
    mdist.mc_sharded_step(run, im0, im5, 101, (v, v + 1))
# FDK
fg = _abi.generic_fdk_geom(45, 65, 33, 40)
proj = torch.from_numpy(np.random.default_rng(5).random((45, 65, 33), dtype=np.float32)).to(dev)
filt = torch.zeros(api.fdk_filtered_shape(fg), device=dev)
z_lo, z_hi = mdist.split_range(fg.nz, ws, rank)
slab = torch.empty((z_hi - z_lo, fg.ny, fg.nx), device=dev)
mdist.fdk_sharded(lambda a, b: api.fdk_filter_dev(fg, proj, filt, a, b, pad=False), lambda: api.fdk_pad_dev(fg, filt),
                  lambda a, b: api.fdk_backproject_dev(fg, filt, slab, a, b), filt, fg.n_views, fg.nv, fg.nz)
slab2 = torch.empty_like(slab)
filt.zero_()
mdist.fdk_sharded_pipelined(lambda a, b: api.fdk_filter_dev(fg, proj, filt, a, b, pad=False),
                            lambda a, b: api.fdk_pad_views_dev(fg, filt, a, b),
                            lambda z0, z1, a, b, cont: api.fdk_backproject_views_dev(fg, filt, slab2, z0, z1, a, b, cont),
                            filt, fg.n_views, fg.nv, fg.nz)
torch.cuda.synchronize()
assert torch.equal(slab, slab2), "pipelined exchange differs from gather-then-backproject"
# z-slabs of equal work (uneven thickness): same voxels
zr = mdist.balanced_split(mdist.fdk_slice_cost(fg), ws, 8)
slab3 = torch.empty((zr[rank][1] - zr[rank][0], fg.ny, fg.nx), device=dev)
filt.zero_()
mdist.fdk_sharded_pipelined(lambda a, b: api.fdk_filter_dev(fg, proj, filt, a, b, pad=False),
                            lambda a, b: api.fdk_pad_views_dev(fg, filt, a, b),
                            lambda z0, z1, a, b, cont: api.fdk_backproject_views_dev(fg, filt, slab3, z0, z1, a, b, cont),
                            filt, fg.n_views, fg.nv, fg.nz, z_ranges=zr)
torch.cuda.synchronize()
np.savez(os.path.join(out, "u%%d.npz" %% rank), slab=slab3.cpu().numpy(), z=np.array(zr[rank]))
# band-limited exchange: one all_to_all of the detector rows each slab reads, rows nobody needs stay NaN
filt.fill_(float("nan"))
slab4 = torch.empty_like(slab)
mdist.fdk_sharded_band(lambda a, b: api.fdk_filter_dev(fg, proj, filt, a, b, pad=False), lambda: api.fdk_pad_dev(fg, filt),
                       lambda a, b: api.fdk_backproject_dev(fg, filt, slab4, a, b), lambda a, b: api.fdk_slab_rows(fg, a, b),
                       filt, fg.n_views, fg.nv, [mdist.split_range(fg.nz, ws, r) for r in range(ws)])
torch.cuda.synchronize()
assert torch.equal(slab, slab4), "band-limited exchange differs from gather-then-backproject"
assert bool(torch.isnan(filt[: fg.n_views * fg.nv]).any()), "every row travelled: the band is not limiting anything"
# >>> peers
# no collective on the data path: the band is loaded out of the peers' buffers (CUDA IPC) by the backprojector's pair conversion
filt.fill_(float("nan"))
slab5 = torch.empty_like(slab)
peers = mdist.PeerRows(api, filt, fg.n_views)
for rep in range(2):                               # twice: the closing fence must keep the rows in place for slow peers
    mdist.fdk_sharded_peers(lambda a, b: api.fdk_filter_dev(fg, proj, filt, a, b, pad=False),
                            lambda z0, z1, ptrs, ends: api.fdk_backproject_peers_dev(fg, ptrs, ends, slab5, z0, z1),
                            peers, fg.n_views, [mdist.split_range(fg.nz, ws, r) for r in range(ws)])
torch.cuda.synchronize()
assert torch.equal(slab, slab5), "peer-memory gather differs from gather-then-backproject"
v_lo, v_hi = mdist.split_range(fg.n_views, ws, rank)
own = filt[v_lo * fg.nv: v_hi * fg.nv]
assert not bool(torch.isnan(own[:, :fg.nu]).any()) and bool(torch.isnan(filt[: fg.n_views * fg.nv]).any()), "only the own views are ever written"
dist.barrier()
peers.close()
# <<< peers
np.savez(os.path.join(out, "r%%d.npz" %% rank), im0=im0.cpu().numpy(), im5=im5.cpu().numpy(), slab=slab2.cpu().numpy(), z=np.array([z_lo, z_hi]))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpus_equal_one(monte, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from monte_b200 import _abi, scenes
    script = os.path.join(str(tmp_path), "w.py")
    with open(script, "w") as f:
        f.write(WORKER % ROOT)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611", script, str(tmp_path)], timeout=600)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(2)]
    lab = scenes.cylinder_phantom(33, 1.0)
    g = scenes.mc_geom(17, 32.5 / 17, n_views=3)
    g.angle_step_deg = 120.0
    r0, r5, _ = monte.simulate(g, scenes.volume_for(lab, 1.0), lab, scenes.make_xs(), scenes.mono_spectrum(), 101, seed=9)
    assert np.array_equal(parts[0]["im0"], r0) and np.array_equal(parts[0]["im5"], r5)
    fg = _abi.generic_fdk_geom(45, 65, 33, 40)
    proj = np.random.default_rng(5).random((45, 65, 33), dtype=np.float32)
    _, vol, _, _ = monte.fdk(fg, proj, want_filtered=False)
    for p in parts:
        assert np.array_equal(p["slab"], vol[p["z"][0]:p["z"][1]])
    for r in range(2):                           # the equal-work partition reconstructs the same voxels
        u = np.load(os.path.join(str(tmp_path), "u%d.npz" % r))
        assert np.array_equal(u["slab"], vol[u["z"][0]:u["z"][1]])
