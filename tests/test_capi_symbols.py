"""CPU checks of the drop-in boundary: the shared library loads without a GPU and exports every
symbol include/afl_b200.h declares; the ctypes table covers the header one to one; argument
validation works without touching a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "afl_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from attacking_federate_learning_b200 import _native
    return _native.lib()


def test_header_lists_expected_entry_points():
    names = header_functions()
    for must in ["afl_sqdist_partial", "afl_krum_select", "afl_bulyan_select", "afl_trimmed_mean", "afl_alie",
                 "afl_mean", "afl_defend_host", "afl_momentum_step"]:
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(lib._name)
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in include/afl_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from attacking_federate_learning_b200 import _native
    assert sorted(_native.SIGNATURES) == header_functions()


def test_argument_validation_without_gpu(lib):
    from attacking_federate_learning_b200 import _native as nat
    assert lib.afl_version().startswith(b"afl_b200")
    assert lib.afl_sqdist_workspace_bytes(100, 11_200_000, nat.AFL_F32, 0) > 100 * 100 * 8
    assert lib.afl_select_workspace_bytes(1000) >= 1000 * 1000 * 8
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.afl_mean(None, 10, 10, 10, 0, None, None) == nat.AFL_ERR_BAD_ARG
    assert lib.afl_trimmed_mean(None, 10, 10, 10, 0, None, 10, 2, None, None) == nat.AFL_ERR_BAD_ARG
    assert lib.afl_bulyan_select(ctypes.c_void_p(16), 10, 10, 2, ctypes.c_void_p(16), None, 0, None) == nat.AFL_ERR_PRECONDITION
    with pytest.raises(AssertionError):
        nat.check(nat.AFL_ERR_PRECONDITION)
    assert lib.afl_defend_host(b"Nope", ctypes.c_void_p(16), 3, 3, 3, 3, 0, None, None, 0) == nat.AFL_ERR_BAD_ARG
    assert b"unknown rule" in lib.afl_last_error()
    assert lib.afl_defend_host(b"Krum", ctypes.c_void_p(16), 10, 8, 8, 10, 5, None, None, 0) == nat.AFL_ERR_PRECONDITION
    assert lib.afl_defend_host(b"Bulyan", ctypes.c_void_p(16), 10, 8, 8, 10, 2, ctypes.c_void_p(16), None, 0) == nat.AFL_ERR_PRECONDITION
