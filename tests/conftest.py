import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the native library if the tree is fresh (nvcc cross-compiles without a GPU)
    so = os.path.join(ROOT, "simuverse_b200", "_native", "liblbm_b200.so")
    if not os.path.exists(so):
        subprocess.check_call([os.path.join(ROOT, "simuverse_b200", "csrc", "build.sh")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # test processes use the product's wire structs on both sides (tests/oracle.py picks them up when the package
    # is already imported; on its own it loads wire.py without the package, see oracle._wire)
    import simuverse_b200.wire  # noqa: F401


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle
