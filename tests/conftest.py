import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_GEN_CACHE = {}


def generated(name, **kw):
    """problem generation is deterministic; cache per test session"""
    from tenscalc_b200 import problems
    key = (name, tuple(sorted(kw.items())))
    if key not in _GEN_CACHE:
        _GEN_CACHE[key] = problems.BUILDERS[name](**kw)
    return _GEN_CACHE[key]
