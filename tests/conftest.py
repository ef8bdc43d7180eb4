import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The shared libraries are build artefacts (git-ignored).  The product loader refuses to run
    # without them; the test session builds them once if a fresh checkout has none.
    from bri17_b200 import _lib
    if not (os.path.exists(_lib.LIB_PATH) and os.path.exists(_lib.RS_LIB_PATH)):
        from bri17_b200 import build
        build.build()


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure): compiled on first use."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    return z, meta
