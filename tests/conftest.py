import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts present (built in-tree by __graft_entry__.build())."""
    from net_ensembles_b200 import _build
    if not (os.path.exists(_build.LIB_PATH) and os.path.exists(_build.HOST_LIB_PATH) and os.path.exists(_build.ORACLE_PATH)):
        _build.build_all()
    return _build


@pytest.fixture(scope="session")
def oracle(built):
    from tests.helpers import Oracle
    return Oracle(built.ORACLE_PATH)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "fixtures.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def ctx(built):
    import net_ensembles_b200 as ne
    c = ne.Context()
    yield c
    c.close()
