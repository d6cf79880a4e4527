import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built libraries are git-ignored: on a fresh checkout build them before any test imports them
    # (nvcc cross-compiles without a GPU; a no-op when they are up to date).
    lib = os.path.join(ROOT, "probabilistic-self-update-line-vector-set-based-point-cloud-registration_b200",
                       "libpsulvsb_b200.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "oracle", "libpsulvsb_oracle.so")):
        import __graft_entry__

        __graft_entry__.build()


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
