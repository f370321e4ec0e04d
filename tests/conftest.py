import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def seed_of(*parts):
    """Reproducible RNG seed from the test's parameters (Python's hash() of strings changes per process)."""
    import zlib
    return zlib.crc32(repr(parts).encode()) & 0xFFFF


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure: make sure the restatement (and, in the authoring container,
    oracle/_ref) is compiled before any test needs it."""
    import oracle
    oracle.build(want_ref=True)
