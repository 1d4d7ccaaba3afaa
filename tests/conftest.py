import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    import pctsea_parent_b200 as P
    from pctsea_parent_b200 import build as B
    B.build()
    c = P.Context(0)  # raises without a GPU: there is no CPU fallback to fall back to
    yield c
    c.close()


@pytest.fixture(scope="session")
def tiny():
    from pctsea_parent_b200 import synth
    return synth.make_dataset(synth.CONFIGS["tiny"], "cpu").numpy() | {"cfg": synth.CONFIGS["tiny"]}


@pytest.fixture(scope="session")
def small():
    from pctsea_parent_b200 import synth
    return synth.make_dataset(synth.CONFIGS["small"], "cpu").numpy() | {"cfg": synth.CONFIGS["small"]}
