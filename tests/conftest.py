import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def glove(oracle):
    """test-data fixture of the reference (1000x50 base, 100x50 queries), parsed to exact f32."""
    store = oracle.load_glove(os.path.join(GOLDEN, "store.txt"))
    queries = oracle.load_glove(os.path.join(GOLDEN, "queries.txt"))
    return store, queries


@pytest.fixture(scope="session")
def glove_index(oracle, glove):
    store, _ = glove
    return oracle.Index(12, None, store.shape[1]).insert_bulk(store)
