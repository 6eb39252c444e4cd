import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED_RNG = 0x45_78_93_F4_4A_B0_67_F0  # the reference's test seed, src/test/mod.rs:19


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must not silently pass on a CPU box: they fail loudly in the fixture instead.
    pass


@pytest.fixture(scope="session")
def oracle_cls():
    from oracle.oracle import Oracle
    return Oracle
