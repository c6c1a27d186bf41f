import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# The library routes batches below BSQ_SMALL_BATCH_READS (default 40000) to its warp-cooperative DP kernels (call latency); the parity
# tests use small batches but are there to check the throughput path (thread-per-extension / thread-per-region kernels), so they switch the
# routing off.  tests/test_gpu_align.py::test_small_batch_routing covers the default routing in a process of its own.
os.environ.setdefault("BSQ_SMALL_BATCH_READS", "0")
# Likewise the finalize stage hands job lists of at most BSQ_FIN_SHORT_LIST (default 1024) regions to its warp-cooperative kernel; the
# parity tests keep the thread-per-region kernels in play.  test_small_batch_routing runs with the defaults.
os.environ.setdefault("BSQ_FIN_SHORT_LIST", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_lib():
    from bioseqdb_b200 import _lib
    L = _lib.lib()
    if L.bsq_device_count() < 1:
        pytest.fail("no CUDA device visible: the -m gpu tests need the B200 box (there is no CPU fallback)")
    return L
