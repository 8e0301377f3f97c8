"""Non-local similarity search: the greedy sub-batch loop and the two operators
the reference gets from the third-party `vpss` package.

Mirrors lib/vnlb/search/search.py: exec_search :25-69, search_and_fill :71-98.
`exec_sim_search_burst` / `fill_patches` keep vpss's call signature, so a module
exposing them is a literal drop-in at search.py:88-89,98."""
import ctypes

import torch

from . import _lib as L
from . import mask as search_mask
from .utils import AttrDict


def exec_sim_search_burst(srch_img, srch_inds, vals, inds, flows, sigma, args, cs_ptr=None):
    """vpss.exec_sim_search_burst (call site search.py:86-89).

    srch_img [T,C,H,W] f32 (YUV); srch_inds i64 [Q,3] (t,y,x); vals f32 [>=Q,k] and
    inds i64 [>=Q,k] are written in place for the first Q rows (ascending by
    distance), rows beyond Q are left as the caller pre-filled them; flows: object
    with .fflow/.bflow [T,2,H,W] or None.  `sigma` is accepted for signature
    compatibility (the l2 search does not use it)."""
    t, c, h, w = srch_img.shape
    q = int(srch_inds.shape[0])
    if q == 0:
        return
    if vals.shape[0] < q or inds.shape[0] < q or vals.shape[1] != args.npatches or inds.shape[1] != args.npatches:
        raise ValueError("vals/inds must be [>=Q, npatches]")
    srch_inds = srch_inds.to(torch.int64).contiguous()
    fflow = getattr(flows, "fflow", None) if flows is not None else None
    bflow = getattr(flows, "bflow", None) if flows is not None else None
    p = L.SearchParams(args.ps, args.pt, args.w_s, args.nWt_f, args.nWt_b, args.npatches,
                       int(args.dist_chnls), args.window_mode_id)
    rc = L.lib.vnlb_search_topk(L.ptr(srch_img, torch.float32), t, c, h, w, L.ptr(srch_inds), q,
                                L.ptr(fflow, torch.float32), L.ptr(bflow, torch.float32), ctypes.byref(p),
                                L.ptr(vals, torch.float32), L.ptr(inds, torch.int64), None, 0,
                                L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_search_topk")


def fill_patches(patches, img, inds, cs_ptr=None):
    """vpss.fill_patches (call site search.py:91-98): gather the pt x c x ps x ps
    patch at every index; rows holding -1 are left untouched."""
    b, k, pt, c, ps, _ = patches.shape
    t, c2, h, w = img.shape
    if c2 != c or tuple(inds.shape) != (b, k):
        raise ValueError("fill_patches: shape mismatch")
    rc = L.lib.vnlb_fill_patches(L.ptr(patches, torch.float32), L.ptr(img, torch.float32),
                                 L.ptr(inds, torch.int64), b, k, t, c, h, w, ps, pt, L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_fill_patches")


def view_batch(tensor, bsize, index):
    """lib/vnlb/utils/batching.py:25-28 as it is effectively called (search.py:49,53):
    rows [index*bsize, (index+1)*bsize)."""
    if tensor is None or not torch.is_tensor(tensor):
        return tensor
    return tensor[index * bsize:(index + 1) * bsize]


def search_and_fill(imgs, patches, bufs, srch_inds, flows, args):
    """search.py:71-98."""
    if args.srch_img == "noisy":
        srch_img = imgs.noisy
    elif args.srch_img == "basic":
        srch_img = imgs.basic
    elif args.srch_img == "clean":
        srch_img = imgs.clean
    else:
        raise ValueError("uknown search image [%s]" % args.srch_img)
    bufs.inds[...] = -1
    bufs.vals[...] = float("inf")
    exec_sim_search_burst(srch_img, srch_inds, bufs.vals, bufs.inds, flows, args.sigma, args)
    for key in imgs.patch_images:
        if imgs[key] is None or patches[key] is None:
            continue
        fill_patches(patches[key], imgs[key], bufs.inds)


def exec_search(patches, imgs, flows, mask, bufs, args):
    """search.py:25-69: `nstreams` sequential sub-batches of `bsize` random reference
    pixels; each sub-batch clears the pixels it found from the mask before the next
    one is drawn (the greedy "paste trick")."""
    bsize = args.bsize
    done = False
    bufs.inds[...] = -1
    bufs.vals[...] = float("inf")
    for index in range(args.nstreams):
        srch_inds = search_mask.mask2inds(mask, bsize)
        if srch_inds.shape[0] == 0:
            done = True
            break
        vbufs = AttrDict({k: view_batch(v, bsize, index) for k, v in bufs.items()})
        vpatches = AttrDict({k: view_batch(v, bsize, index) for k, v in patches.items()})
        search_and_fill(imgs, vpatches, vbufs, srch_inds, flows, args)
        search_mask.update_mask_inds(mask, vbufs.inds, args.c, boost=args.aggreBoost)
    done = done or (mask.sum().item() == 0)
    return done


def exec_refinement(patches, bufs, sigma, thresh=2.0):
    """lib/vnlb/search/refinement.py:15-29 (disabled in the reference at proc_nl.py:70): rows whose
    mean distance ratio vals[:,1:]/vals[:,1] exceeds `thresh` are dropped (all their indices set to -1).
    A few elementwise torch ops on [rows,k]; not on the hot path."""
    vals = bufs.vals
    ave_vals = torch.mean(vals[:, 1:] / vals[:, [1]], 1)
    noupdate = torch.nonzero(ave_vals > thresh)
    bufs.inds[noupdate] = -1
