"""RGB <-> "YUV" (orthonormal opponent transform of the C++ VNLB code).
Mirrors lib/vnlb/utils/color.py: rgb2yuv_images :10-13, yuv2rgb_images :15-18."""
import torch

from . import _lib as L


def rgb2yuv(burst, cs_ptr=None):
    """rgb2yuv_cpp, color.py:52-77: returns a new tensor."""
    t, c, h, w = burst.shape
    out = torch.empty_like(burst)
    L.check(L.lib.vnlb_rgb2yuv(L.ptr(burst, torch.float32), L.ptr(out), t, c, h, w, L.stream_ptr(cs_ptr)),
            "vnlb_rgb2yuv")
    return out


def yuv2rgb(burst, cs_ptr=None):
    """apply_yuv2rgb, color.py:31-50: in place."""
    t, c, h, w = burst.shape
    L.check(L.lib.vnlb_yuv2rgb(L.ptr(burst, torch.float32), L.ptr(burst), t, c, h, w, L.stream_ptr(cs_ptr)),
            "vnlb_yuv2rgb")
    return burst


def yuv2rgb_new(burst, cs_ptr=None):
    """Out-of-place yuv2rgb (the input keeps its YUV values)."""
    t, c, h, w = burst.shape
    out = torch.empty_like(burst)
    L.check(L.lib.vnlb_yuv2rgb(L.ptr(burst, torch.float32), L.ptr(out), t, c, h, w, L.stream_ptr(cs_ptr)),
            "vnlb_yuv2rgb")
    return out


def rgb2yuv_images(images):
    for key in images.ikeys:
        if images[key] is None:
            continue
        images[key] = rgb2yuv(images[key])


def yuv2rgb_images(images):
    for key in images.ikeys:
        if images[key] is None:
            continue
        yuv2rgb(images[key])
