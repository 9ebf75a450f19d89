"""The drop-in boundary: the C-ABI shared library loads and exports every symbol that
include/tenscalc_b200.h declares; misuse is reported through return codes; without a GPU the
library refuses to create a solver (there is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, generated


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tenscalc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[A-Za-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from tenscalc_b200.build import build_library
    from tenscalc_b200.runtime import load_library
    build_library()
    return load_library()


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tenscalc_b200.h but not exported"


def test_python_binding_covers_header(lib):
    from tenscalc_b200.runtime import ABI
    assert sorted(n for n, _, _ in ABI) == _declared_symbols()


def test_error_codes(lib):
    assert lib.tc_version() >= 100
    assert lib.tc_error_string(0) == b"ok"
    assert b"no CPU fallback" in lib.tc_error_string(-3)
    h = ctypes.c_void_p()
    assert lib.tc_create(ctypes.byref(h), None, 0, 1, 0) == -1          # TC_EINVAL
    assert lib.tc_destroy(None) == -1
    assert lib.tc_batch(None) == -1


def test_image_is_validated(lib):
    from tenscalc_b200 import tape
    gen, _ = generated("sls")
    img = tape.serialize(gen)
    h = ctypes.c_void_p()
    ndev = lib.tc_device_count()
    bad = b"\0" * len(img)
    rc = lib.tc_create(ctypes.byref(h), bad, len(bad), 4, 0)
    assert rc == (-2 if ndev > 0 else -3)                                 # TC_EFORMAT / TC_ENOGPU
    if ndev == 0:
        rc = lib.tc_create(ctypes.byref(h), img, len(img), 4, 0)
        assert rc == -3, "without a CUDA device tc_create must fail loudly, not fall back"


def test_no_cpu_fallback_in_python():
    import torch
    from tenscalc_b200.runtime import TenscalcError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    gen, _ = generated("sls")
    with pytest.raises(TenscalcError):
        gen(8)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "tenscalc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"#include\s+\"[^\"]*oracle/", src), f


def test_image_layout_roundtrip():
    """the serialized image header carries the #define table of cmex2optimizeCS.m:303-331"""
    from tenscalc_b200 import tape
    gen, _ = generated("mpc_dcmotor")
    img = tape.serialize(gen)
    ih = np.frombuffer(img[:64 * 8], dtype=np.int64)
    dh = np.frombuffer(img[64 * 8:64 * 8 + 16 * 8], dtype=np.float64)
    assert ih[0] == tape.MAGIC
    assert (ih[16], ih[17], ih[18]) == (89, 60, 174)
    assert ih[1] == gen.problem.n_slots and ih[2] == len(gen.problem.groups)
    assert dh[0] == 1e-4 and dh[3] == 1e-5 and dh[4] == 1e-7 and dh[5] == 1.0
    # the image ends with the names of the set functions and outputs (tc_find_set / tc_find_output), NUL-terminated
    names = [n for n, _ in gen.parameters] + [n for n, _ in gen.variables] + [n for n, _ in gen.outputs]
    blob = b"".join(n.encode() + b"\0" for n in names)
    assert ih[13] == len(blob) and ih[30] == 0                      # allowSave off
    assert img.rstrip(b"\0").endswith(blob.rstrip(b"\0"))
    assert tape.serialize(generated("mpc_dcmotor", allowSave=True)[0])[30 * 8] == 1


@pytest.mark.gpu
def test_name_lookup(lib):
    """`<class>_set_<name>` of the reference (writeCfunctionpergroup.m:524) is tc_set(h, tc_find_set(h, "<name>"), ...)
    here: every parameter, variable and output of the class resolves to its index, unknown names do not, and a
    truncated names section is a format error"""
    from tenscalc_b200 import tape
    gen, _ = generated("mpc_dcmotor")
    img = tape.serialize(gen)
    h = ctypes.c_void_p()
    assert lib.tc_create(ctypes.byref(h), img, len(img), 64, 0) == 0
    sets = [n for n, _ in gen.parameters] + [n for n, _ in gen.variables]
    assert lib.tc_num_sets(h) == len(sets) and lib.tc_num_outputs(h) == len(gen.outputs)
    for k, n in enumerate(sets):
        assert lib.tc_find_set(h, n.encode()) == k and lib.tc_set_name(h, k) == n.encode()
        shape = dict(gen.parameters + gen.variables)[n]
        assert lib.tc_set_size(h, k) == (int(np.prod(shape)) if shape else 1)
    for k, (n, _) in enumerate(gen.outputs):
        assert lib.tc_find_output(h, n.encode()) == k and lib.tc_output_name(h, k) == n.encode()
    assert lib.tc_find_set(h, b"no_such_parameter") == -1 and lib.tc_find_output(h, b"") == -1
    assert lib.tc_set_name(h, len(sets)) is None and lib.tc_find_set(h, None) == -1
    assert lib.tc_destroy(h) == 0
    ih = np.frombuffer(img[:64 * 8], dtype=np.int64).copy()
    ih[13] -= 3                                                     # the last name loses its terminator
    bad = ih.tobytes() + img[64 * 8:]
    assert lib.tc_create(ctypes.byref(h), bad, len(bad), 64, 0) == -2
