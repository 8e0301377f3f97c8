"""ctypes binding of libvnlb_b200.so (the C ABI declared in include/vnlb_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing
this module without the built library raises immediately."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvnlb_b200.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
WINDOW_SHIFT, WINDOW_CLIP = 0, 1
EIG_TRIDIAG, EIG_JACOBI = 0, 1


class VnlbError(RuntimeError):
    pass


class SearchParams(ctypes.Structure):
    _fields_ = [("ps", ctypes.c_int32), ("pt", ctypes.c_int32), ("w_s", ctypes.c_int32),
                ("nWt_f", ctypes.c_int32), ("nWt_b", ctypes.c_int32), ("k", ctypes.c_int32),
                ("dist_chnls", ctypes.c_int32), ("window_mode", ctypes.c_int32)]


class BayesParams(ctypes.Structure):
    _fields_ = [("step", ctypes.c_int32), ("k", ctypes.c_int32), ("ps", ctypes.c_int32),
                ("pt", ctypes.c_int32), ("c", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("sigma2", ctypes.c_float), ("sigmab2", ctypes.c_float), ("thresh", ctypes.c_float),
                ("cov_from_basic", ctypes.c_int32), ("eig_method", ctypes.c_int32)]


EXPORTS = [
    "vnlb_kernel_launches",
    "vnlb_last_error", "vnlb_version", "vnlb_rgb2yuv", "vnlb_yuv2rgb", "vnlb_init_mask", "vnlb_init_mask_tile",
    "vnlb_set_search_path", "vnlb_search_workspace_bytes", "vnlb_search_topk", "vnlb_fill_patches", "vnlb_mask_update",
    "vnlb_count_mask", "vnlb_select_queries", "vnlb_pad_queries", "vnlb_round_dedup", "vnlb_round_draw", "vnlb_round_dedup_dev",
    "vnlb_flat_areas", "vnlb_bayes_workspace_bytes", "vnlb_bayes_filter", "vnlb_bayes_debug", "vnlb_bayes_matrix_dim", "vnlb_bayes_fused_supported", "vnlb_set_bayes_split", "vnlb_set_filter_mma", "vnlb_bayes_aggregate_fused", "vnlb_aggregate",
    "vnlb_normalize",
]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "vnlb_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or ./build.sh).  There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)
lib.vnlb_last_error.restype = ctypes.c_char_p
lib.vnlb_search_workspace_bytes.restype = ctypes.c_size_t
lib.vnlb_bayes_workspace_bytes.restype = ctypes.c_size_t
lib.vnlb_kernel_launches.restype = ctypes.c_ulonglong
for _name in EXPORTS:
    getattr(lib, _name)  # fail at import if a declared symbol is not exported

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
lib.vnlb_rgb2yuv.argtypes = [_vp, _vp, _i, _i, _i, _i, _vp]
lib.vnlb_yuv2rgb.argtypes = [_vp, _vp, _i, _i, _i, _i, _vp]
lib.vnlb_init_mask.argtypes = [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
lib.vnlb_init_mask_tile.argtypes = [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
lib.vnlb_set_search_path.argtypes = [_i]
lib.vnlb_search_workspace_bytes.argtypes = [_i, ctypes.POINTER(SearchParams)]
lib.vnlb_search_topk.argtypes = [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, ctypes.POINTER(SearchParams),
                                 _vp, _vp, _vp, _sz, _vp]
lib.vnlb_fill_patches.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
lib.vnlb_mask_update.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]
lib.vnlb_count_mask.argtypes = [_vp, _i, _i, _i, _vp, _vp]
lib.vnlb_select_queries.argtypes = [_vp, _i, _i, _i, ctypes.c_double, ctypes.c_uint32, ctypes.c_uint32, _vp, _i, _vp, _vp]
lib.vnlb_pad_queries.argtypes = [_vp, _vp, _i, _vp]
lib.vnlb_round_dedup.argtypes = [_vp, _vp, _i, _i, _vp, ctypes.c_uint32, _vp, _i, _i, _i, _i, _i, _vp, _vp]
lib.vnlb_round_draw.argtypes = [_vp, _i, _i, _i, ctypes.c_double, _i, _i, ctypes.c_uint32, _vp, _vp, _vp, _vp]
lib.vnlb_round_dedup_dev.argtypes = [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]
lib.vnlb_flat_areas.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]
lib.vnlb_bayes_workspace_bytes.argtypes = [_i, ctypes.POINTER(BayesParams)]
lib.vnlb_bayes_filter.argtypes = [_vp, _vp, _vp, _vp, _i, ctypes.POINTER(BayesParams), _vp, _vp, _sz, _vp]
lib.vnlb_bayes_debug.argtypes = [_vp, _vp, _vp, _vp, _i, ctypes.POINTER(BayesParams), _vp, _vp, _vp, _vp, _vp, _sz, _vp]
lib.vnlb_bayes_matrix_dim.argtypes = [ctypes.POINTER(BayesParams), ctypes.POINTER(ctypes.c_int)]
lib.vnlb_bayes_fused_supported.argtypes = [ctypes.POINTER(BayesParams)]
lib.vnlb_set_bayes_split.argtypes = [_i]
lib.vnlb_set_filter_mma.argtypes = [_i]
lib.vnlb_bayes_aggregate_fused.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.POINTER(BayesParams), _f, _vp, _vp, _vp, _sz, _vp]
lib.vnlb_aggregate.argtypes = [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]
lib.vnlb_normalize.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _vp]


launches = 0   # C-ABI calls that enqueued one of our kernels (bench.py reports it)


class StageTimer:
    """Optional per-stage device timing (CUDA events on the launching stream).
    Installed by bench.py; `None` in normal operation."""

    def __init__(self):
        self.events = {}

    def start(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return (name, ev)

    def stop(self, tok):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.setdefault(tok[0], []).append((tok[1], ev))

    def summary(self):
        torch.cuda.synchronize()
        return {k: dict(ms=sum(a.elapsed_time(b) for a, b in v), launches=len(v)) for k, v in self.events.items()}


timer = None


on_launch = None   # optional callable(launch_count): bench.py samples NVML clocks from it while the GPU is busy


def check(rc, what):
    global launches
    launches += 1
    if on_launch is not None:
        on_launch(launches)
    if rc != OK:
        msg = lib.vnlb_last_error().decode()
        if rc == ERR_BAD_ARG:
            raise ValueError("%s: %s" % (what, msg))
        raise VnlbError("%s failed (%d): %s" % (what, rc, msg))


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("vnlb_b200 operates on CUDA tensors only (got a %s tensor)" % t.device)
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise ValueError("expected dtype %s, got %s" % (dtype, t.dtype))
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(cs_ptr=None):
    if cs_ptr is None:
        cs_ptr = torch.cuda.current_stream().cuda_stream
    return ctypes.c_void_p(cs_ptr)
