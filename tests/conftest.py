import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fixture_index():
    import whippet_jl_b200 as wq
    return wq.FlatIndex.load(os.path.join(GOLDEN, "test_index_flat.npz"))


@pytest.fixture(scope="session")
def test_param():
    """AlignParam of test/runtests.jl:283-284."""
    import whippet_jl_b200 as wq
    return wq.AlignParam(1, 2, 4, 4, 4, 5, 1, 2, 1000, 0.05, 0.7, False, False, True, False, True)
