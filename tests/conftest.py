import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_bytes(name: str) -> bytes:
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def libzstd():
    """The real libzstd 1.5.7 (present in this image on the build container and the GPU box)."""
    from oracle.libzstd_ref import load
    try:
        return load("1.5.7")
    except OSError:
        pytest.skip("no libzstd shared library in this environment")


@pytest.fixture(scope="session")
def kzg_lib():
    """libkzg.so built in-tree (nvcc cross-compiles without a GPU)."""
    from kompressor_b200 import build as kb
    from kompressor_b200._lib import lib
    kb.build()
    return lib()
