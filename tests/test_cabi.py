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


def test_round_draw_rejects_bad_arguments_without_a_gpu():
    from vnlb_b200 import _lib
    n0 = int(_lib.lib.vnlb_kernel_launches())
    rc = _lib.lib.vnlb_round_draw(ctypes.c_void_p(8), 2, 8, 8, 0.0, 16, 64, 1, ctypes.c_void_p(8), ctypes.c_void_p(8),
                                  ctypes.c_void_p(8), None)
    assert rc == _lib.ERR_BAD_ARG and b"vnlb_round_draw" in _lib.lib.vnlb_last_error()
    rc = _lib.lib.vnlb_round_dedup_dev(ctypes.c_void_p(8), ctypes.c_void_p(8), 4, 4, ctypes.c_void_p(8), None,
                                       ctypes.c_void_p(8), 2, 3, 8, 8, 1, ctypes.c_void_p(8), None)
    assert rc == _lib.ERR_BAD_ARG
    assert int(_lib.lib.vnlb_kernel_launches()) == n0
    prev = _lib.lib.vnlb_set_search_path(1)
    assert prev in (0, 1, 2) and _lib.lib.vnlb_set_search_path(prev) == 1
    prev = _lib.lib.vnlb_set_filter_mma(1)
    assert prev in (0, 1) and _lib.lib.vnlb_set_filter_mma(prev) == 1


def test_switches_and_counters_without_a_gpu():
    """vnlb_set_bayes_split returns the previous setting; vnlb_kernel_launches is a monotonic counter (0 launches here)."""
    from vnlb_b200 import _lib
    prev = _lib.lib.vnlb_set_bayes_split(0)
    assert prev in (0, 1)
    assert _lib.lib.vnlb_set_bayes_split(1) == 0
    assert _lib.lib.vnlb_set_bayes_split(prev) == 1
    n0 = int(_lib.lib.vnlb_kernel_launches())
    _lib.lib.vnlb_rgb2yuv(None, None, 1, 3, 4, 4, None)      # rejected before any launch
    assert int(_lib.lib.vnlb_kernel_launches()) == n0


def test_bayes_workspace_contract_without_a_gpu():
    """vnlb_bayes_workspace_bytes reports the split path's real size (the library never allocates: ADVICE r1) and
    a call without the workspace is refused with VNLB_ERR_WORKSPACE before any launch."""
    from vnlb_b200 import _lib
    p1 = _lib.BayesParams(0, 100, 7, 2, 3, 39, 400., 400., 2.7, 0, _lib.EIG_TRIDIAG)
    p2 = _lib.BayesParams(1, 60, 7, 2, 3, 39, 400., 0., 0.7, 1, _lib.EIG_TRIDIAG)
    prev = _lib.lib.vnlb_set_bayes_split(1)
    try:
        b1 = int(_lib.lib.vnlb_bayes_workspace_bytes(1, ctypes.byref(p1)))
        b2 = int(_lib.lib.vnlb_bayes_workspace_bytes(1, ctypes.byref(p2)))
        assert 30_000 * 3 < b1 < 45_000 * 3 and 10_000 * 3 < b2 < 30_000 * 3          # ~41 KB / ~27 KB per (group, channel)
        assert int(_lib.lib.vnlb_bayes_workspace_bytes(1000, ctypes.byref(p1))) == 1000 * b1
        assert int(_lib.lib.vnlb_bayes_workspace_bytes(10 ** 6, ctypes.byref(p1))) == 16384 * b1   # chunked beyond a round
        p3 = _lib.BayesParams(0, 40, 5, 1, 3, 20, 400., 400., 2.7, 0, _lib.EIG_TRIDIAG)           # other shapes: single kernel
        assert int(_lib.lib.vnlb_bayes_workspace_bytes(64, ctypes.byref(p3))) == 0
        n0 = int(_lib.lib.vnlb_kernel_launches())
        rc = _lib.lib.vnlb_bayes_filter(ctypes.c_void_p(16), None, None, None, 4, ctypes.byref(p1), None, None, 0, None)
        assert rc == _lib.ERR_WORKSPACE and b"workspace" in _lib.lib.vnlb_last_error()
        rc = _lib.lib.vnlb_bayes_aggregate_fused(ctypes.c_void_p(16), None, ctypes.c_void_p(16), 4, 4, 3, 32, 32,
                                                 ctypes.byref(p1), 0., ctypes.c_void_p(16), ctypes.c_void_p(16),
                                                 ctypes.c_void_p(16), b1 - 1, None)
        assert rc == _lib.ERR_WORKSPACE
        assert int(_lib.lib.vnlb_kernel_launches()) == n0
        _lib.lib.vnlb_set_bayes_split(0)
        assert int(_lib.lib.vnlb_bayes_workspace_bytes(64, ctypes.byref(p1))) == 0
    finally:
        _lib.lib.vnlb_set_bayes_split(prev)


def test_library_never_allocates_or_synchronises():
    """The header's contract, checked on the sources: no cudaMalloc / cudaFree / *Synchronize in the product library."""
    import glob
    for f in glob.glob(os.path.join(ROOT, "vnlb_b200", "csrc", "*.cu*")):
        src = re.sub(r"//.*", "", open(f).read())
        for bad in ("cudaMalloc", "cudaFree", "cudaDeviceSynchronize", "cudaStreamSynchronize"):
            assert bad not in src, (f, bad)
