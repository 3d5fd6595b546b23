import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "supernova_b200", "libsupernova_b200.so")
    if not os.path.exists(so) or not os.path.exists(os.path.join(ROOT, "supernova_b200", "sn_build_graph")):
        g.build()
    return so
