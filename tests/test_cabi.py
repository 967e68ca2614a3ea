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


def _integration_md_stub():
    """The ctypes stub INTEGRATION.md §1 shows a TLS maintainer, executed verbatim (only the library path is
    made absolute), and the reference's own arguments for it (ragged lc_arr, structured overview)."""
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes, numpy\n.*?)```", text, flags=re.S).group(1)
    block = block.replace('ctypes.CDLL("libtlsb200.so")', "ctypes.CDLL(%r)" % native.library_path())
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    g = load_search_golden("tiny")
    tp = g["templates"]
    lc_arr = np.empty(len(tp["length"]), dtype=object)
    for r in range(len(lc_arr)):
        lc_arr[r] = np.array(tp["signal"][tp["offset"][r]: tp["offset"][r] + tp["length"][r]])
    overview = np.zeros(len(lc_arr), dtype=[("duration", "f8"), ("width_in_samples", "i8"), ("overshoot", "f8")])
    overview["width_in_samples"], overview["overshoot"] = tp["width"], tp["overshoot"]
    return g, (lambda: ns["search_periods"](g["periods"], g["t"], g["y"], g["dy"], lc_arr, overview, **g["params"]))


def test_integration_md_stub_refuses_without_a_gpu(has_cuda):
    """On a machine without a GPU the stub must marshal its arguments and surface the library's refusal as a
    RuntimeError (there is no CPU fallback)."""
    if has_cuda:
        pytest.skip("a CUDA device is present: test_integration_md_stub_runs_as_written covers this machine")
    _, call = _integration_md_stub()
    with pytest.raises(RuntimeError, match="CUDA"):
        call()


@pytest.mark.gpu
def test_integration_md_stub_runs_as_written():
    """With a GPU the stub, exactly as INTEGRATION.md prints it, returns what tls_b200.native returns and what the
    reference's numba path returned for the same arguments (the golden)."""
    g, call = _integration_md_stub()
    chi2, row, depth = call()
    want = native.search_periods(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"])
    np.testing.assert_array_equal(chi2, want[0])
    np.testing.assert_array_equal(row, want[1])
    np.testing.assert_array_equal(depth, want[2])
    from conftest import assert_search_parity

    assert_search_parity((chi2, row, depth), g, rtol=1e-5, label="INTEGRATION.md stub")
