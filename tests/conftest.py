import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref_ops():
    """Reference CPU backend (oracle/_ref) as the checker; None if not built."""
    import unpaper_gpu_b200 as U
    from oracle import checker
    lib = checker.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libunpaper_ref.so not built")
    return U.HostOps(lib, "ref_host_")


@pytest.fixture(scope="session")
def ref_lib():
    import unpaper_gpu_b200 as U
    from oracle import checker
    lib = checker.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libunpaper_ref.so not built")
    return lib


@pytest.fixture(scope="session")
def cuda_ops(ref_ops):
    if os.environ.get("UNPAPER_TEST_SELFCHECK") == "ref":
        # harness self-check on a CPU-only box: compare the checker with itself
        return ref_ops
    from unpaper_gpu_b200 import lib as L
    handle = L.load()
    if handle.unpaper_cuda_try_init() != 0:
        pytest.fail("CUDA backend unavailable on a GPU run (no fallback exists)")
    return L.host_ops()
