import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ref_golden():
    return dict(np.load(os.path.join(GOLDEN, "ref_golden.npz")))


@pytest.fixture(scope="session")
def edge_golden():
    return dict(np.load(os.path.join(GOLDEN, "edge_golden.npz")))


@pytest.fixture(scope="session")
def layer_golden():
    return dict(np.load(os.path.join(GOLDEN, "layer_golden.npz")))


@pytest.fixture(scope="session")
def bench_positions():
    return dict(np.load(os.path.join(GOLDEN, "bench_positions.npz")))


@pytest.fixture(scope="session")
def oracle_nets():
    from leela_b200 import synth
    from oracle import oracle
    return oracle.OracleNet(synth.policy_weights()), oracle.OracleNet(synth.value_weights())
