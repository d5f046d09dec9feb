import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def relerr(a, b):
    """norm-wise relative error |a-b| / |b| (the measure BASELINE.json's 1e-5 gate is stated in)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    from oracle.make_golden import scene_from_dict
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    return d, scene_from_dict(d, name)
