import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def alego():
    import alego_pkg
    mod = alego_pkg.load()
    if not (os.path.exists(mod.LIB_PATH) and os.path.exists(mod.SYNTH_PATH)):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def ob():
    from oracle import binding
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    return binding


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.argwhere(a != b)
    if len(d) == 0:
        return "equal"
    i = tuple(d[0])
    return "%d diffs, first at %s: %s vs %s" % (len(d), i, a[i], b[i])
