"""The C-ABI library loads and exports every symbol include/vnlb_b200.h declares
(CPU only: no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vnlb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vnlb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    so = os.path.join(ROOT, "vnlb_b200", "libvnlb_b200.so")
    assert os.path.exists(so), "build the extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s


def test_binding_lists_every_header_symbol():
    from vnlb_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib.vnlb_version() >= 100


def test_struct_layouts_match_header():
    from vnlb_b200 import _lib
    assert ctypes.sizeof(_lib.SearchParams) == 8 * 4
    assert ctypes.sizeof(_lib.BayesParams) == 11 * 4


def test_bad_arguments_are_rejected_without_a_gpu():
    from vnlb_b200 import _lib
    rc = _lib.lib.vnlb_rgb2yuv(None, None, 1, 3, 4, 4, None)
    assert rc == _lib.ERR_BAD_ARG
    assert b"vnlb_rgb2yuv" in _lib.lib.vnlb_last_error()
    p = _lib.SearchParams(7, 2, 26, 6, 6, 100, 1, 0)          # even window
    rc = _lib.lib.vnlb_search_topk(ctypes.c_void_p(8), 4, 3, 32, 32, ctypes.c_void_p(8), 1, None, None,
                                   ctypes.byref(p), ctypes.c_void_p(8), ctypes.c_void_p(8), None, 0, None)
    assert rc == _lib.ERR_BAD_ARG


def test_switches_and_counters_without_a_gpu():
    """vnlb_set_bayes_split returns the previous setting; vnlb_kernel_launches is a monotonic counter (0 launches here)."""
    from vnlb_b200 import _lib
    prev = _lib.lib.vnlb_set_bayes_split(0)
    assert prev in (0, 1, 2)
    assert _lib.lib.vnlb_set_bayes_split(1) == 0
    assert _lib.lib.vnlb_set_bayes_split(prev) == 1
    n0 = int(_lib.lib.vnlb_kernel_launches())
    _lib.lib.vnlb_rgb2yuv(None, None, 1, 3, 4, 4, None)      # rejected before any launch
    assert int(_lib.lib.vnlb_kernel_launches()) == n0
