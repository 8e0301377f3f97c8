"""Weighted patch aggregation.  Mirrors lib/vnlb/agg/comp_agg.py: agg_patches
:47-60 -> compute_agg_batch :62-79 -> exec_agg_simple_numba :106-138 (the
reference's live path: single CPU thread behind whole-video copies)."""
import torch

from . import _lib as L


def compute_agg_batch(deno, patches, inds, weights, vals, ivals, ps, ps_t, cs_ptr=None):
    """comp_agg.py:62-79; `vals` / `ivals` are accepted and unused, as in the
    reference (:140-141)."""
    b, k = inds.shape
    t, c, h, w = deno.shape
    L.check(L.lib.vnlb_aggregate(L.ptr(patches, torch.float32), L.ptr(inds, torch.int64), b, k,
                                 L.ptr(deno, torch.float32), L.ptr(weights, torch.float32), t, c, h, w, ps, ps_t,
                                 L.stream_ptr(cs_ptr)), "vnlb_aggregate")


def agg_patches(patches, images, bufs, args, cs_ptr=None):
    """comp_agg.py:47-60.  Invalid rows are skipped on the device."""
    return compute_agg_batch(images.deno, patches.noisy, bufs.inds, images.weights, bufs.vals, images.vals,
                             args.ps, args.ps_t, cs_ptr)


def normalize(images, args, cs_ptr=None):
    """lib/vnlb/proc_nl.py:118-125: deno /= weights where weights != 0, elsewhere
    deno = basic (second step) or noisy (first step)."""
    t, c, h, w = images.deno.shape
    fill = images.basic if args.step == 1 else images.noisy
    L.check(L.lib.vnlb_normalize(L.ptr(images.deno, torch.float32), L.ptr(images.weights, torch.float32),
                                 L.ptr(fill, torch.float32), t, c, h, w, L.stream_ptr(cs_ptr)), "vnlb_normalize")
