"""Per-group Bayes estimate.  Mirrors lib/vnlb/deno/__init__.py:4-10 (dispatch)
and lib/vnlb/deno/bayes_est.py:17-62 (denoise)."""
import ctypes
import threading

import torch

from . import _lib as L

# Workspace of the split Bayes path (include/vnlb_b200.h: vnlb_bayes_workspace_bytes): caller-owned, so it comes from
# torch's caching allocator -- one grow-only buffer per (device, stream), allocated while that stream is current (the
# allocator then recycles it in stream order).  The library itself never allocates.
_ws_lock = threading.Lock()
_ws_cache = {}


def bayes_workspace(p, b, device, cs_ptr=None):
    """(tensor or None, nbytes) for a call of `b` groups with parameters `p` on the current stream."""
    nbytes = int(L.lib.vnlb_bayes_workspace_bytes(int(b), ctypes.byref(p)))
    if nbytes == 0:
        return None, 0
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           int(cs_ptr) if cs_ptr is not None else int(torch.cuda.current_stream(dev).cuda_stream))
    with _ws_lock:
        ws = _ws_cache.get(key)
        if ws is None or ws.numel() < nbytes:
            if len(_ws_cache) >= 16 and ws is None:
                _ws_cache.pop(next(iter(_ws_cache)))        # oldest (device, stream) entry
            _ws_cache[key] = ws = torch.empty((nbytes + nbytes // 8,), dtype=torch.uint8, device=dev)
    return ws, int(ws.numel())


def release_workspaces():
    """Drop the cached Bayes workspaces (they return to torch's caching allocator)."""
    with _ws_lock:
        _ws_cache.clear()


def bayes_denoise(patches, args, inds=None, cs_ptr=None):
    """bayes_est.denoise: filters patches.noisy [b,n,pt,c,ps,ps] in place;
    patches.basic is read only (the reference re-centres it back to its input,
    bayes_est.py:52).  Rows not valid in `inds` are skipped on the device (the
    reference compacts them on the host: proc_nl.py:160-177).  Returns rank_var [b]."""
    b, n, pt, c, ps, _ = patches.noisy.shape
    p = L.BayesParams(int(args.step), n, ps, pt, c, int(args.rank), float(args.sigma2), float(args.sigmab2),
                      float(args.thresh), int(args.cpatches == "basic"), args.eig_method_id)
    rank_var = torch.empty((b,), dtype=torch.float32, device=patches.noisy.device)
    ws, nbytes = bayes_workspace(p, b, patches.noisy.device, cs_ptr)
    flat = patches.get("flat")
    rc = L.lib.vnlb_bayes_filter(L.ptr(patches.noisy, torch.float32), L.ptr(patches.basic, torch.float32),
                                 L.ptr(flat, torch.uint8), L.ptr(inds, torch.int64), b, ctypes.byref(p), L.ptr(rank_var),
                                 L.ptr(ws), nbytes, L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_bayes_filter")
    return rank_var


def bayes_debug(patches, args, inds=None, cs_ptr=None):
    """Parity hook (vnlb_bayes_debug): bayes_denoise that also returns what compute_cov_mat / denoise_eigvals /
    bayes_filter_coeff (bayes_est.py:112-144) produce: dict(mat [b,c,q,q], is_gram, lam [b,c,40], coef [b,c,40], m [b,c]).
    `mat` is the covariance (q = p) or, for k + 8 <= p (step 2), the Gram matrix Y Y^T / n (q = k)."""
    b, n, pt, c, ps, _ = patches.noisy.shape
    p = bayes_params(args, n, c)
    gram = ctypes.c_int(0)
    q = int(L.lib.vnlb_bayes_matrix_dim(ctypes.byref(p), ctypes.byref(gram)))
    if q == 0:
        raise ValueError("bayes_debug: patch shape outside the tridiagonal kernels' envelope")
    dev = patches.noisy.device
    mat = torch.zeros((b, c, q, q), dtype=torch.float32, device=dev)
    lam = torch.zeros((b, c, 40), dtype=torch.float32, device=dev)
    coef = torch.zeros((b, c, 40), dtype=torch.float32, device=dev)
    m = torch.zeros((b, c), dtype=torch.int32, device=dev)
    ws, nbytes = bayes_workspace(p, b, dev, cs_ptr)
    rc = L.lib.vnlb_bayes_debug(L.ptr(patches.noisy, torch.float32), L.ptr(patches.basic, torch.float32),
                                L.ptr(patches.get("flat"), torch.uint8), L.ptr(inds, torch.int64), b, ctypes.byref(p),
                                L.ptr(mat), L.ptr(lam), L.ptr(coef), L.ptr(m), L.ptr(ws), nbytes, L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_bayes_debug")
    return dict(mat=mat, is_gram=bool(gram.value), lam=lam, coef=coef, m=m)


def bayes_params(args, n, c):
    return L.BayesParams(int(args.step), int(n), int(args.ps), int(args.pt), int(c), int(args.rank),
                         float(args.sigma2), float(args.sigmab2), float(args.thresh),
                         int(args.cpatches == "basic"), args.eig_method_id)


def fused_supported(args, c, shape=None):
    """True iff the fused gather+Bayes+aggregate kernel serves these parameters (and, when `shape` = (T,C,H,W) is
    given, a video of that size: the kernel keeps 32-bit image offsets)."""
    p = bayes_params(args, args.npatches, c)
    if shape is not None and int(shape[0]) * int(shape[1]) * int(shape[2]) * int(shape[3]) >= 2 ** 31:
        return False
    return args.eig_method == "tridiag" and bool(L.lib.vnlb_bayes_fused_supported(ctypes.byref(p)))


def bayes_aggregate_fused(images, inds, args, cs_ptr=None):
    """fill_patches + flat areas + bayes_est.denoise + agg_patches in one kernel
    (the throughput schedule's inner step): gathers the groups named by `inds`
    from images.noisy / images.basic (YUV) and adds the filtered patches into
    images.deno / images.weights."""
    t, c, h, w = images.noisy.shape
    b, k = inds.shape
    p = bayes_params(args, k, c)
    ws, nbytes = bayes_workspace(p, b, images.noisy.device, cs_ptr)
    rc = L.lib.vnlb_bayes_aggregate_fused(L.ptr(images.noisy, torch.float32), L.ptr(images.basic, torch.float32),
                                          L.ptr(inds, torch.int64), b, t, c, h, w, ctypes.byref(p),
                                          float(args.gamma * args.sigma2), L.ptr(images.deno, torch.float32),
                                          L.ptr(images.weights, torch.float32), L.ptr(ws), nbytes,
                                          L.stream_ptr(cs_ptr))
    L.check(rc, "vnlb_bayes_aggregate_fused")


def denoise(patches, params, method="bayes", inds=None):
    if method == "bayes":
        return bayes_denoise(patches, params, inds)
    raise ValueError("Uknown denoising method [%s]" % method)
