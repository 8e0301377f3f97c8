"""Flat-group flag (second step only).  Mirrors lib/vnlb/utils/flat_areas.py:
update_flat_patch :8-14, exec_flat_areas :16-34."""
import torch

from . import _lib as L


def exec_flat_areas(flat_patch, patches, gamma, sigma2, inds=None, cs_ptr=None):
    b, k, pt, c, ps, _ = patches.shape
    L.check(L.lib.vnlb_flat_areas(L.ptr(patches, torch.float32), L.ptr(inds, torch.int64),
                                  L.ptr(flat_patch, torch.uint8), b, k, c, ps, pt, float(gamma * sigma2),
                                  L.stream_ptr(cs_ptr)), "vnlb_flat_areas")


def update_flat_patch(patches, args, inds=None):
    if args.step == 1:
        exec_flat_areas(patches.flat, patches.noisy, args.gamma, args.sigma2, inds)
