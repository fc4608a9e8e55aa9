import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


@pytest.fixture(scope="session")
def oracle():
    """oracle/liboracle.so built on demand (gcc only)."""
    from oracle import binding
    binding.build(ref=True)
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def monte():
    """libmonte_gpu bound to cuda:0.  No fallback: raises if the library or the GPU is missing."""
    from monte_b200 import api
    api.init(int(os.environ.get("LOCAL_RANK", "0")))
    yield api
    api.shutdown()


@pytest.fixture(scope="session")
def monte_emu():
    """The monte_b200.api surface bound to tests/emu/_build/libmonte_gpu_emu.so: the library's own .cu
    sources compiled by g++ against a SIMT emulation (tests/emu/cuda_runtime.h).  Test infrastructure
    only -- it lets the CPU suite run kernel and host logic; it is not a fallback of the product."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("monte_emu_build", os.path.join(ROOT, "tests", "emu", "build.py"))
    eb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eb)
    api = eb.api()
    api.init(0)
    yield api
    api.shutdown()


GOLDEN = os.path.join(ROOT, "tests", "golden")
