import os
import sys

import pytest

# before anything initialises CUDA: enough hardware queues for groups whose members SHARE a device (tests/test_gpu_group.py),
# where a spinning exchange kernel must never sit in front of the other member's kernels (see engine_one.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native():
    """The built CUDA library; building is the job of __graft_entry__.build()."""
    import kzg_poly_commit_exploration_b200 as pkg

    if not os.path.exists(pkg.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return pkg


@pytest.fixture(scope="session")
def ctx(native):
    c = native.Context(0)
    yield c
    c.close()
