import gc
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ctx():
    import gkr_msm_b200 as g
    c = g.Context(0)  # raises loudly when there is no GPU / no built extension: no CPU fallback
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _collect_between_tests():
    """device objects are released by __del__; objects caught in reference cycles (ctx <-> object) wait for the cyclic
    collector, and a context has a bounded number of result slots -- collect after every test so a long parametrised
    run (or one slowed down by compute-sanitizer) never runs out"""
    yield
    gc.collect()
