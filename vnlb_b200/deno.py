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


def denoise(patches, params, method="bayes", inds=None):
    if method == "bayes":
        return bayes_denoise(patches, params, inds)
    raise ValueError("Uknown denoising method [%s]" % method)
