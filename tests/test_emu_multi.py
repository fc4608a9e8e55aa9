"""CPU: the multi-device mode of the C ABI (monte_gpu_init(ndev > 1)) under SIMT emulation with MONTE_EMU_DEVICES
"devices" -- per-device contexts, photon / view / z-slab partition, the peer-memory tally reduce fused with the
counts -> map epilogue, the peer-memory row-band gather of the backprojector, and the NCCL call pattern (against
the in-process stand-in of tests/emu/emu_runtime.cpp).  Bodies: tests/test_multi_gpu.py.  Test infrastructure only."""
import os

import numpy as np
import pytest

import test_multi_gpu as M
from monte_b200 import _abi


@pytest.fixture()
def emu8(monte_emu):
    os.environ["MONTE_EMU_DEVICES"] = "8"
    try:
        yield monte_emu
    finally:
        monte_emu.init(0)
        del os.environ["MONTE_EMU_DEVICES"]


@pytest.mark.parametrize("n_dev,case", [(2, M.FDK_CASES[0]), (3, M.FDK_CASES[1]), (2, M.FDK_CASES[2]), (4, M.FDK_CASES[3]), (8, M.FDK_CASES[0])])
def test_emu_fdk_multi_equals_single(emu8, n_dev, case):
    M.body_fdk_multi_equals_single(emu8, n_dev, case)


@pytest.mark.parametrize("n_dev,chunks,case", [(2, 4, (37, 40, 30, 48, None)), (3, 4, (37, 40, 30, 48, "roi")), (8, 3, (5, 24, 16, 16, None)),
                                               (8, 4, (130, 24, 20, 24, None)), (2, 4, (64, 300, 12, 16, None)), (5, 2, (12, 33, 65, 24, "tall")),
                                               (4, 4, (9, 16, 5, 16, None)), (2, 8, (37, 40, 30, 48, None)), (4, 8, (130, 24, 20, 24, "roi")),
                                               (3, 7, (20, 24, 20, 24, None))])
def test_emu_fdk_multi_chunked_pipeline(emu8, n_dev, chunks, case):
    M.body_fdk_multi_chunked(emu8, n_dev, chunks, case)


@pytest.mark.parametrize("n_dev,mode", [(2, "p2p"), (3, "nccl"), (8, "p2p"), (5, "nccl")])
def test_emu_mc_multi_equals_single(emu8, n_dev, mode):
    M.body_mc_multi_equals_single(emu8, n_dev, mode)


def test_emu_more_devices_than_photons_or_slices(emu8):
    """devices that get an empty photon range / no z-slice / no view must be harmless"""
    M.body_mc_multi_equals_single(emu8, 8, "p2p", per=11)            # n_range (5, 8): 3 photons for 8 devices
    g = _abi.generic_fdk_geom(5, 24, 16, 8)                           # 5 views, 8 slices, 8 devices
    proj = np.random.default_rng(1).random((5, 24, 16), dtype=np.float32)
    emu8.init(0)
    f1, v1, _, _ = emu8.fdk(g, proj)
    emu8.init(list(range(8)))
    f2, v2, _, _ = emu8.fdk(g, proj)
    assert np.array_equal(f1, f2) and np.array_equal(v1, v2)


def test_emu_without_peer_access(emu8):
    """no NVLink P2P between the devices: MC falls back to the NCCL reduce, FDK refuses loudly"""
    os.environ["MONTE_EMU_NO_PEER"] = "1"
    try:
        emu8.init([0, 1])
        assert emu8.load().monte_gpu_peer_access() == 0
        M.body_mc_multi_equals_single(emu8, 2, "p2p")                 # asks for p2p, gets nccl
        emu8.init([0, 1])
        g = _abi.generic_fdk_geom(6, 24, 16, 16)
        with pytest.raises(emu8.MonteError, match="peer access"):
            emu8.fdk(g, np.zeros((6, 24, 16), np.float32))
    finally:
        del os.environ["MONTE_EMU_NO_PEER"]
        emu8.init(0)


@pytest.mark.parametrize("n_dev", [2, 5])
def test_emu_hu_volume_and_ring_detector_multi(emu8, n_dev):
    M.body_hu_volume_multi_equals_single(emu8, n_dev)


def test_emu_label_cache(monte_emu):
    M.body_label_cache(monte_emu)


@pytest.mark.parametrize("n_dev", [2, 3, 8])
def test_emu_scattered_label_upload(emu8, n_dev):
    M.body_scattered_label_upload(emu8, n_dev)


def test_emu_argument_errors(emu8):
    M.body_argument_errors(emu8)
