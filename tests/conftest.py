import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "baseline", "_ref")):
    if p not in sys.path and os.path.isdir(p):
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def have_exauq():
    try:
        import exauq.core.modelling  # noqa: F401

        return True
    except ImportError:
        return False


needs_exauq = pytest.mark.skipif(not have_exauq(), reason="reference package 'exauq' not importable")
