import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def rank_dict(g, r):
    """Fields of rank r of a golden fixture as a plain dict."""
    pre = "r%d_" % r
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
