"""Per-group Bayes estimate.  Mirrors lib/vnlb/deno/__init__.py:4-10 (dispatch)
and lib/vnlb/deno/bayes_est.py:17-62 (denoise)."""
import ctypes

import torch

from . import _lib as L


def bayes_denoise(patches, args, inds=None, cs_ptr=None):
    """bayes_est.denoise: filters patches.noisy [b,n,pt,c,ps,ps] in place;
    patches.basic is read only (the reference re-centres it back to its input,
    bayes_est.py:52).  Rows not valid in `inds` are skipped on the device (the
    reference compacts them on the host: proc_nl.py:160-177).  Returns rank_var [b]."""
    b, n, pt, c, ps, _ = patches.noisy.shape
    p = L.BayesParams(int(args.step), n, ps, pt, c, int(args.rank), float(args.sigma2), float(args.sigmab2),
                      float(args.thresh), int(args.cpatches == "basic"), args.eig_method_id)
    rank_var = torch.empty((b,), dtype=torch.float32, device=patches.noisy.device)
    nbytes = L.lib.vnlb_bayes_workspace_bytes(b, ctypes.byref(p))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=patches.noisy.device)
    flat = patches.get("flat")
    rc = L.lib.vnlb_bayes_filter(L.ptr(patches.noisy, torch.float32), L.ptr(patches.basic, torch.float32),
                                 L.ptr(flat, torch.uint8), L.ptr(inds, torch.int64), b, ctypes.byref(p), L.ptr(rank_var),
                                 L.ptr(ws), nbytes, L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_bayes_filter")
    return rank_var


def bayes_params(args, n, c):
    return L.BayesParams(int(args.step), int(n), int(args.ps), int(args.pt), int(c), int(args.rank),
                         float(args.sigma2), float(args.sigmab2), float(args.thresh),
                         int(args.cpatches == "basic"), args.eig_method_id)


def fused_supported(args, c):
    p = bayes_params(args, args.npatches, c)
    return args.eig_method == "tridiag" and bool(L.lib.vnlb_bayes_fused_supported(ctypes.byref(p)))


def bayes_aggregate_fused(images, inds, args, cs_ptr=None):
    """fill_patches + flat areas + bayes_est.denoise + agg_patches in one kernel
    (the throughput schedule's inner step): gathers the groups named by `inds`
    from images.noisy / images.basic (YUV) and adds the filtered patches into
    images.deno / images.weights."""
    t, c, h, w = images.noisy.shape
    b, k = inds.shape
    p = bayes_params(args, k, c)
    rc = L.lib.vnlb_bayes_aggregate_fused(L.ptr(images.noisy, torch.float32), L.ptr(images.basic, torch.float32),
                                          L.ptr(inds, torch.int64), b, t, c, h, w, ctypes.byref(p),
                                          float(args.gamma * args.sigma2), L.ptr(images.deno, torch.float32),
                                          L.ptr(images.weights, torch.float32), L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_bayes_aggregate_fused")


def denoise(patches, params, method="bayes", inds=None):
    if method == "bayes":
        return bayes_denoise(patches, params, inds)
    raise ValueError("Uknown denoising method [%s]" % method)
