"""pytest configuration: registers the `gpu` marker and the shared fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "kats.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def lib_vectors():
    return np.load(os.path.join(GOLDEN, "lib_vectors.npz"))


@pytest.fixture(scope="session")
def cli_manifest():
    with open(os.path.join(GOLDEN, "cli", "manifest.json")) as fh:
        return json.load(fh)


