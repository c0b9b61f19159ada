import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "xiangqi-alphazero_b200")
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import xq_oracle
    xq_oracle.lib()
    return xq_oracle


@pytest.fixture(scope="session")
def rules_golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "rules_golden.npz")))


@pytest.fixture(scope="session")
def attacked_golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "attacked_golden.npz")))


@pytest.fixture(scope="session")
def mcts_golden():
    import json
    with open(os.path.join(GOLDEN, "mcts_golden.json")) as f:
        return json.load(f)
