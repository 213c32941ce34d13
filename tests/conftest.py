import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # test infrastructure is built on demand; the product library is built by __graft_entry__.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emul")])


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """A live context on cuda:0 through the C ABI.  No fallback: fails if the library or GPU is missing."""
    from coherence_renderer_b200 import abi

    c = abi.Context(0)
    yield c
    c.close()
