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
    # make sure the C-ABI library exists (nvcc cross-compiles without a GPU); a stale or missing
    # .so must never make the suite fall back to anything else
    try:
        import importlib.util

        spec = importlib.util.spec_from_file_location("exq_build", os.path.join(ROOT, "exauq-toolbox_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if mod.needs_build() and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            mod.build()
    except Exception as exc:  # the tests that need the library will say so
        print(f"[conftest] could not build libexauq_b200.so: {exc}", file=sys.stderr)


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
