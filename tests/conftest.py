import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_2k():
    from oracle_lib import Oracle
    o = Oracle(2048, 1024)
    yield o
    o.close()


@pytest.fixture(scope="session")
def gpu_2k():
    import vvc_extension_mm_b200 as mm
    g = mm.MmGpu(2048, 1024, max_refs=4)
    yield g
    g.close()
