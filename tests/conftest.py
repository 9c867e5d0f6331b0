import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def summaries():
    with open(os.path.join(GOLDEN, "summaries.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def checkpoint_index():
    with open(os.path.join(GOLDEN, "checkpoint_index.json")) as f:
        return json.load(f)


@pytest.fixture(autouse=True)
def _fresh_block_counters():
    """The reference keeps process-global class counters (model.py:326,401); every test starts
    from the fresh-process state."""
    try:
        from x3d_tf_b200 import model as _m
        _m.reset_block_counters()
    except Exception:
        pass
    yield
