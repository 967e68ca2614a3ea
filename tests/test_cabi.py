"""The C-ABI shared library: it loads without a GPU, exports every symbol that
include/tlsb200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO, load_search_golden
from tls_b200 import native

HEADER = os.path.join(REPO, "include", "tlsb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tlsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = native.lib()
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(L, name), "libtlsb200.so does not export " + name
    assert sorted(native.EXPORTS) == names, "tls_b200/native.py EXPORTS out of sync with include/tlsb200.h"


def test_library_is_in_tree_and_versioned():
    assert os.path.dirname(native.library_path()) == os.path.join(REPO, "tls_b200")
    assert b"sm_100a" in native.lib().tlsb_version()


def test_struct_layouts_match_the_header():
    # plain pointers and sizes only: 3 pointers + int64, 5 pointers + int64, 6 doubles, pointer + int32
    assert ctypes.sizeof(native.LightCurve) == 32
    assert ctypes.sizeof(native.Templates) == 48
    assert ctypes.sizeof(native.Params) == 48
    assert ctypes.sizeof(native.Exec) == 16


def test_no_cpu_fallback(has_cuda):
    if has_cuda:
        pytest.skip("a CUDA device is present; the refusal path is for machines without one")
    g = load_search_golden("tiny")
    assert native.device_count() == 0
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        native.search_periods(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"])
    with pytest.raises(RuntimeError):
        native.Searcher()


def test_null_arguments_are_rejected_without_touching_cuda():
    L = native.lib()
    rc = L.tlsb_search_periods(None, None, 0, None, None, None, None, None, None, None)
    assert rc == -1  # TLSB_ERR_ARG
    assert b"NULL" in L.tlsb_last_error()
    assert L.tlsb_set_plan_mode(None, 0) == -1
    assert L.tlsb_destroy(None) == 0
    assert L.tlsb_last_launch_count(None) == 0
