import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


MEASUREMENT_CASES = [
    # (name, subsystem_dims, measurement_subsystems, memory_slot_indices, num_memory_slots, max_outcome_level)
    ("two_transmons_both", [3, 3], [0, 1], [0, 1], None, 1),
    ("two_transmons_q1_slot2", [3, 3], [1], [2], 3, None),
    ("two_transmons_q0_levels", [3, 3], [0], [0], None, 2),
    ("mixed_dims_swapped_slots", [2, 3, 2], [0, 2], [1, 0], None, 1),
]  # same table as tests/golden/make_golden.py


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def max_col_l2(a, b):
    """max over columns of the L2 error -- the north-star parity metric (bar: < 1e-8)."""
    a = np.asarray(a)
    b = np.asarray(b)
    d = a - b
    if d.ndim == 1:
        return float(np.linalg.norm(d))
    return float(np.max(np.linalg.norm(d.reshape(d.shape[0], -1) if d.ndim == 2 else d.reshape(-1, d.shape[-1]), axis=0)))
