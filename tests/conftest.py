import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def micro_path():
    from openhush_b200 import synth
    return synth.model_path("micro")


@pytest.fixture(scope="session")
def micro_oracle(micro_path):
    from oracle import oracle as O
    m = O.OracleModel(micro_path)
    yield m
    m.close()


@pytest.fixture(scope="session")
def clip1():
    from openhush_b200 import synth
    return synth.audio(1)
