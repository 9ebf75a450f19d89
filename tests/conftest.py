import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need the native library AND a visible CUDA device: skip them elsewhere, so that a
    plain `pytest tests` is green on a CPU-only machine (the product itself still fails loudly:
    BatchedSolver raises without a device, tests/test_abi.py)"""
    reason = None
    try:
        from tenscalc_b200.runtime import load_library
        if load_library().tc_device_count() == 0:
            reason = "no CUDA device visible"
    except Exception as e:          # library not built
        reason = f"native library unavailable: {e}"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


_GEN_CACHE = {}


def generated(name, **kw):
    """problem generation is deterministic; cache per test session"""
    from tenscalc_b200 import problems
    key = (name, tuple(sorted(kw.items())))
    if key not in _GEN_CACHE:
        _GEN_CACHE[key] = problems.BUILDERS[name](**kw)
    return _GEN_CACHE[key]
