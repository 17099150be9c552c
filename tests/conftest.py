import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_report_header(config):
    """Which libpagegeom.so the run exercises (PAGEGEOM_LIB selects a variant build, e.g. the checked one)."""
    try:
        from multimodal_embeddings_b200 import _lib
        path = os.environ.get("PAGEGEOM_LIB") or _lib.LIB_PATH
        return f"libpagegeom: {path} (checked build: {bool(_lib.lib().pg_build_checked())})"
    except Exception as e:  # not built yet: the tests themselves say so
        return f"libpagegeom: not loadable ({e})"


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".json.gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    if name.endswith(".json"):
        with open(path) as f:
            return json.load(f)
    if name.endswith(".npz"):
        return np.load(path)
    raise ValueError(name)


@pytest.fixture(scope="session")
def f1_pages():
    return load_golden("f1_pages.json.gz")


@pytest.fixture(scope="session")
def f4():
    return load_golden("f4_stage45.json")
