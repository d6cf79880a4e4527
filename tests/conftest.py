import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    gd = os.path.join(ROOT, "tests", "golden")
    reg = dict(np.load(os.path.join(gd, "registration_test.npz")))
    bench = dict(np.load(os.path.join(gd, "benchmark.npz")))
    with open(os.path.join(gd, "golden.json")) as f:
        meta = json.load(f)
    return {"reg": reg, "bench": bench, "meta": meta}
