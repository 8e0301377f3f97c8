"""
oracle/vnlb_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + the C library built from oracle/vnlb_oracle.c) of the
reference's `vnlb.denoise` hot path, one function per reference function, each
citing the reference file:line it follows (paths relative to /root/reference).

Pinning status
--------------
* Bayes / flat areas / aggregation / mask / colour / normalise / orchestration:
  PINNED against the reference's own code, run in the build container on CPU
  through import shims (tests/golden/make_golden.py); the outputs are the
  committed fixtures under tests/golden/ and tests/test_oracle_golden.py checks
  this file against them.
* Similarity search + patch gather (`vpss`, third-party, un-pinned, absent):
  PARITY UNPINNED -- restated from the published algorithm and the call-site
  contract; see the header of oracle/vnlb_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (vnlb_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# ----------------------------------------------------------------------------
# C library (search, gather, aggregation)
# ----------------------------------------------------------------------------

class _SearchParams(ctypes.Structure):
    _fields_ = [("ps", ctypes.c_int), ("pt", ctypes.c_int), ("w_s", ctypes.c_int),
                ("nWt_f", ctypes.c_int), ("nWt_b", ctypes.c_int), ("k", ctypes.c_int),
                ("dist_chnls", ctypes.c_int), ("window_mode", ctypes.c_int)]


def build(force=False):
    """Compile oracle/vnlb_oracle.c -> oracle/liboracle.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "vnlb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_search_topk.restype = ctypes.c_int
        _LIB.oracle_search_all.restype = ctypes.c_int
        _LIB.oracle_fill_patches.restype = ctypes.c_int
        _LIB.oracle_aggregate.restype = ctypes.c_int
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _search_params(args, dist_chnls=None, window_mode="shift"):
    if dist_chnls is None:
        # luminance-only distance in the first step, all channels in the second
        # (published C++ VNLB; `isFirstStep` is carried by lib/vnlb/params.py:19
        # for exactly this and read nowhere in-tree)
        dist_chnls = 1 if args.step == 0 else args.c
    return _SearchParams(args.ps, args.pt, args.w_s, args.nWt_f, args.nWt_b, args.npatches,
                         int(dist_chnls), 0 if window_mode == "shift" else 1)


def exec_sim_search_burst(srch_img, srch_inds, vals, inds, flows, sigma, args,
                          dist_chnls=None, window_mode="shift"):
    """vpss.exec_sim_search_burst restated; call site lib/vnlb/search/search.py:86-89.
    Writes vals [Q..,k] f32 and inds [Q..,k] i64 in place (rows >= len(srch_inds)
    keep the caller's sentinels)."""
    img = np.ascontiguousarray(srch_img, dtype=np.float32)
    T, C, H, W = img.shape
    q = np.ascontiguousarray(srch_inds, dtype=np.int64).reshape(-1, 3)
    Q = q.shape[0]
    assert vals.dtype == np.float32 and inds.dtype == np.int64
    assert vals.flags.c_contiguous and inds.flags.c_contiguous
    assert vals.shape[0] >= Q and vals.shape[1] == args.npatches
    ff = bf = None
    if flows is not None:
        ff = np.ascontiguousarray(flows["fflow"], dtype=np.float32)
        bf = np.ascontiguousarray(flows["bflow"], dtype=np.float32)
        if not ff.any() and not bf.any():
            ff = bf = None
    p = _search_params(args, dist_chnls, window_mode)
    rc = lib().oracle_search_topk(_fp(img), T, C, H, W, _fp(q), Q, _fp(ff), _fp(bf),
                                  ctypes.byref(p), _fp(vals), _fp(inds))
    if rc != 0:
        raise ValueError("oracle_search_topk: bad arguments")


def search_all(srch_img, t0, y0, x0, flows, args, dist_chnls=None, window_mode="shift"):
    """All candidate (distance, index) pairs of one query in enumeration order."""
    img = np.ascontiguousarray(srch_img, dtype=np.float32)
    T, C, H, W = img.shape
    ff = bf = None
    if flows is not None:
        ff = np.ascontiguousarray(flows["fflow"], dtype=np.float32)
        bf = np.ascontiguousarray(flows["bflow"], dtype=np.float32)
    p = _search_params(args, dist_chnls, window_mode)
    nmax = (args.nWt_f + args.nWt_b + 1) * args.w_s * args.w_s
    d = np.empty(nmax, np.float32)
    ind = np.empty(nmax, np.int64)
    n = lib().oracle_search_all(_fp(img), T, C, H, W, int(t0), int(y0), int(x0), _fp(ff), _fp(bf),
                                ctypes.byref(p), _fp(d), _fp(ind))
    return d[:n], ind[:n]


def fill_patches(patches, img, inds, ps=None, pt=None):
    """vpss.fill_patches restated; call site lib/vnlb/search/search.py:91-98."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    T, C, H, W = img.shape
    B, K = inds.shape
    _, _, pt_, C_, ps_, _ = patches.shape
    assert patches.dtype == np.float32 and patches.flags.c_contiguous and C_ == C
    lib().oracle_fill_patches(_fp(patches), _fp(img), _fp(np.ascontiguousarray(inds)),
                              B, K, T, C, H, W, ps_, pt_)


# ----------------------------------------------------------------------------
# parameters  (lib/vnlb/params.py)
# ----------------------------------------------------------------------------

def default_params(sigma):
    """lib/vnlb/params.py:11-50 (the classic VNLB values; pairs = [step1, step2])."""
    return dict(
        aggreBoost=[True, True], bsize=[128, 128], c=[3, 3], flatAreas=[False, True],
        gamma=[0.95, 0.2], nSimilarPatches=[100, 60], nstreams=[8, 18], procStep=[3, 3],
        rank=[39, 39], sigma=[sigma, sigma], sigmaBasic=[sigma, 0.], sizePatch=[7, 7],
        sizePatchTime=[2, 2], sizeSearchTimeBwd=[6, 6], sizeSearchTimeFwd=[6, 6],
        sizeSearchWindow=[27, 27], srch_img=["noisy", "basic"], cpatches=["noisy", "basic"],
        variThres=[2.7, 0.7])


def get_args(params, c, step):
    """lib/vnlb/params.py:102-214: per-step view with the reference's shortcuts."""
    a = SimpleNamespace(**{k: v[step] for k, v in params.items()})
    a.c, a.step = c, step
    a.ps, a.pt = a.sizePatch, a.sizePatchTime
    a.npatches = a.nSimilarPatches
    a.w_s, a.nWt_f, a.nWt_b = a.sizeSearchWindow, a.sizeSearchTimeFwd, a.sizeSearchTimeBwd
    a.sigma2, a.sigmab2 = a.sigma ** 2, a.sigmaBasic ** 2
    a.thresh = a.variThres
    a.tsize = a.nstreams * a.bsize                      # params.py:222-233
    return a


# ----------------------------------------------------------------------------
# colour  (lib/vnlb/utils/color.py)
# ----------------------------------------------------------------------------

def rgb2yuv(burst):
    """rgb2yuv_cpp, lib/vnlb/utils/color.py:52-77 (returns a new array)."""
    b = np.asarray(burst, dtype=np.float32)
    w0 = np.float32(1. / np.sqrt(3))
    w1 = np.float32(1. / np.sqrt(2))
    w2 = np.float32(np.sqrt(2.) * 2. / np.sqrt(3))
    q, h = np.float32(.25), np.float32(.5)
    out = np.empty_like(b)
    out[:, 0] = w0 * ((b[:, 0] + b[:, 1]) + b[:, 2])
    out[:, 1] = w1 * (b[:, 0] - b[:, 2])
    out[:, 2] = w2 * ((q * b[:, 0] - h * b[:, 1]) + q * b[:, 2])
    return out


def yuv2rgb(burst):
    """apply_yuv2rgb, lib/vnlb/utils/color.py:31-50 (returns a new array)."""
    b = np.asarray(burst, dtype=np.float32)
    w0 = np.float32(1. / np.sqrt(3))
    w1 = np.float32(1. / np.sqrt(2))
    w2 = np.float32(np.sqrt(2.) / np.sqrt(3))
    w2h = np.float32(np.sqrt(2.) / np.sqrt(3) * 0.5)
    y, u, v = b[:, 0], b[:, 1], b[:, 2]
    out = np.empty_like(b)
    out[:, 0] = (w0 * y + w1 * u) + w2h * v
    out[:, 1] = w0 * y - w2 * v
    out[:, 2] = (w0 * y - w1 * u) + w2h * v
    return out


# ----------------------------------------------------------------------------
# reference-pixel mask  (lib/vnlb/search_mask/mask.py)
# ----------------------------------------------------------------------------

def init_mask(shape, args):
    """init_mask -> comp_params -> fill_mask, mask.py:190-213,252-288,315-358
    (no partition borders: origin 0, ending = full size)."""
    t, c, h, w = shape
    step = args.procStep
    end_t, end_h, end_w = t - args.pt + 1, h - args.ps + 1, w - args.ps + 1
    mask = np.zeros((t, h, w), np.int8)
    ti = np.arange(end_t)[:, None, None]
    hi = np.arange(end_h)[None, :, None]
    wi = np.arange(end_w)[None, None, :]
    last_t = ti == end_t - 1
    phase_h = np.where(last_t, 0, ti)                       # step_t == 1 (mask.py:248)
    take_h = (hi % step) == (phase_h % step)
    first_h, last_h = hi == 0, hi == end_h - 1
    row_ok = take_h | first_h | last_h
    phase_w = np.where(last_h, 0, phase_h + hi // step)
    take_w = (wi % step) == (phase_w % step)
    first_w, last_w = wi == 0, wi == end_w - 1
    sel = row_ok & (take_w | first_w | last_w)
    mask[:end_t, :end_h, :end_w] = sel
    return mask, int(sel.sum())


def mask2inds(mask, bsize, randperm):
    """mask.py:18-31: `randperm(N)` must return the reference's
    th.randperm(N) draw (CPU default generator)."""
    index = np.argwhere(mask)                                # row-major (t,y,x) like th.nonzero
    if index.shape[0] == 0:
        return index
    order = np.asarray(randperm(index.shape[0]))
    return index[order[:bsize]]


def update_mask_inds(mask, inds, c, boost=True):
    """mask.py:37-86 + agg_boost / agg_boost_cuda :104-187: clear every found
    neighbour and (aggregation boost) its 4 spatial neighbours."""
    t, h, w = mask.shape
    hw, chw = h * w, c * h * w
    if inds.shape[0] == 0:
        return
    inds = inds[np.all(inds != -1, 1)]
    if inds.shape[0] == 0:
        return
    ti = (inds // chw).ravel()
    hi = ((inds % hw) // w).ravel()
    wi = (inds % w).ravel()
    deltas = [(0, 0, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0), (0, -1, 0)] if boost else [(0, 0, 0)]
    valid_ind = (ti >= 0) & (ti < t) & (hi >= 0) & (hi < h) & (wi >= 0) & (wi < w)
    for dt, dh, dw in deltas:
        mt, mh, mw = ti + dt, hi + dh, wi + dw
        ok = valid_ind & (mt >= 0) & (mt < t) & (mh >= 0) & (mh < h) & (mw >= 0) & (mw < w)
        mask[mt[ok], mh[ok], mw[ok]] = 0


def exec_refinement(vals, inds, thresh=2.0):
    """lib/vnlb/search/refinement.py:15-29 (in place on inds)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ave = np.mean(vals[:, 1:] / vals[:, [1]], 1)
    inds[ave > thresh] = -1


# ----------------------------------------------------------------------------
# flat areas  (lib/vnlb/utils/flat_areas.py)
# ----------------------------------------------------------------------------

def exec_flat_areas(pnoisy, gamma, sigma2):
    """flat_areas.py:16-34: unbiased variance over all k*p samples per channel,
    mean over channels, flat = var < gamma*sigma^2."""
    b, n, pt, c, ph, pw = pnoisy.shape
    pflat = pnoisy.transpose(0, 3, 1, 2, 4, 5).reshape(b, c, -1)
    Z = pflat.shape[2]
    psum = pflat.sum(2, dtype=np.float32)
    psum2 = (pflat ** 2).sum(2, dtype=np.float32)
    var = (psum2 - (psum * psum / np.float32(Z))) / np.float32(Z - 1)
    var = var.mean(1, dtype=np.float32)
    return var < np.float32(gamma * sigma2)


# ----------------------------------------------------------------------------
# Bayes estimate  (lib/vnlb/deno/bayes_est.py)
# ----------------------------------------------------------------------------

_EIGH_THREADS = 1


def set_num_threads(n):
    """Host threads of the batched eigendecomposition (the dominant CPU cost; bench.py's CPU arm sets it to the core
    count -- numpy's batched eigh is a serial loop over LAPACK calls that releases the GIL, so the batch is split
    over a thread pool; per-matrix results are unchanged).  The C search uses OpenMP (OMP_NUM_THREADS)."""
    global _EIGH_THREADS
    _EIGH_THREADS = max(1, int(n))


def _eigh_batched(cov):
    """np.linalg.eigh over the batch.  98 x 98 problems are far below the size where a threaded BLAS pays: OpenBLAS'
    own threads are limited to 1 inside (measured here: 10.2 s -> 4.3 s for 3072 matrices) and the batch is split
    over `_EIGH_THREADS` host threads instead (-> 1.6 s on 8 vCPUs)."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:                                   # pragma: no cover
        from contextlib import nullcontext as threadpool_limits
    nt = min(_EIGH_THREADS, cov.shape[0] // 8)
    with threadpool_limits(limits=1):
        if nt <= 1:
            return np.linalg.eigh(cov)
        from concurrent.futures import ThreadPoolExecutor
        chunks = np.array_split(np.arange(cov.shape[0]), nt)
        with ThreadPoolExecutor(nt) as ex:
            parts = list(ex.map(lambda ix: np.linalg.eigh(cov[ix[0]:ix[-1] + 1]), chunks))
    return np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts], 0)


def bayes_denoise(pnoisy, pbasic, flat, args, return_parts=False):
    """bayes_est.denoise, lib/vnlb/deno/bayes_est.py:17-62.

    pnoisy, pbasic: [b,n,pt,c,ph,pw] f32; flat: [b] bool.  Returns the filtered
    noisy stack, the (re-centred) basic stack and rank_var [b]; with
    return_parts also the covariance, eigenvalues and filter coefficients."""
    b, n, pt, c, ph, pw = pnoisy.shape
    step2 = args.step == 1
    # flat_pdim: 'b n pt c ph pw -> b c n (pt ph pw)'            :25,69-71
    X = np.ascontiguousarray(pnoisy.transpose(0, 3, 1, 2, 4, 5)).reshape(b, c, n, -1).astype(np.float32)
    B = np.ascontiguousarray(pbasic.transpose(0, 3, 1, 2, 4, 5)).reshape(b, c, n, -1).astype(np.float32)
    # center_patches                                             :88-110
    cbasic = None
    if step2:
        cbasic = B.mean(2, keepdims=True, dtype=np.float32)
        B = B - cbasic
    cnoisy = X.mean(2, keepdims=True, dtype=np.float32)
    if step2:
        fl = np.flatnonzero(np.asarray(flat))
        cnoisy[fl] = cbasic[fl]
    X = X - cnoisy
    # flat_bdim + compute_cov_mat                                :32-36,112-126
    Xf = X.reshape(b * c, n, -1)
    Bf = B.reshape(b * c, n, -1)
    pin = Xf if args.cpatches == "noisy" else Bf
    cov = np.matmul(pin.transpose(0, 2, 1), pin) / np.float32(n)
    evals, evecs = _eigh_batched(cov)
    evals = evals[:, ::-1].astype(np.float32).copy()
    evecs = evecs[:, :, ::-1][:, :, :args.rank].astype(np.float32)
    rank_var = evals.reshape(b, c, -1).sum(2).mean(1)            # :39-40
    # denoise_eigvals ("clipped")                                :129-138
    lam = evals.copy()
    lam[:, :args.rank] -= np.minimum(lam[:, :args.rank], np.float32(args.sigmab2))
    # bayes_filter_coeff                                         :140-144
    s2 = np.float32(args.sigma2)
    thr = np.float32(args.thresh * args.sigma2)
    with np.errstate(divide="ignore", invalid="ignore"):
        coeff = np.where(lam > thr, np.float32(1.) / (np.float32(1.) + s2 / lam), np.float32(0.))
    coeff = coeff.astype(np.float32)
    # filter_patches                                             :146-151
    Z = np.matmul(Xf, evecs)
    R = evecs * coeff[:, None, :args.rank]
    Xf = np.matmul(Z, R.transpose(0, 2, 1))
    # expand + re-centre                                         :48-55
    X = Xf.reshape(b, c, n, -1) + cnoisy
    if step2:
        B = B + cbasic
    out_n = X.reshape(b, c, n, pt, ph, pw).transpose(0, 2, 3, 1, 4, 5)
    out_b = B.reshape(b, c, n, pt, ph, pw).transpose(0, 2, 3, 1, 4, 5)
    out_n = np.ascontiguousarray(out_n, dtype=np.float32)
    out_b = np.ascontiguousarray(out_b, dtype=np.float32)
    if return_parts:
        return out_n, out_b, rank_var, dict(cov=cov, evals=evals, coeff=coeff, evecs=evecs)
    return out_n, out_b, rank_var


# ----------------------------------------------------------------------------
# aggregation / normalisation  (lib/vnlb/agg/comp_agg.py, lib/vnlb/proc_nl.py)
# ----------------------------------------------------------------------------

def agg_patches(deno, weights, pnoisy, inds):
    """agg_patches -> exec_agg_simple_numba, comp_agg.py:47-60,106-138
    (in place on deno [T,C,H,W] and weights [T,H,W])."""
    valid = np.all(inds != -1, 1)
    vp = np.ascontiguousarray(pnoisy[valid], dtype=np.float32)
    vi = np.ascontiguousarray(inds[valid], dtype=np.int64)
    T, C, H, W = deno.shape
    B, K = vi.shape
    _, _, pt, _, ps, _ = vp.shape
    assert deno.dtype == np.float32 and weights.dtype == np.float32
    lib().oracle_aggregate(_fp(deno), _fp(weights), _fp(vp), _fp(vi), B, K, T, C, H, W, ps, pt)


def normalize(deno, weights, fill_img):
    """proc_nl.py:118-125: deno/=weights where weights != 0, else fill image."""
    w = np.repeat(weights[:, None], deno.shape[1], 1)
    nz = w != 0
    deno[nz] /= w[nz]
    deno[~nz] = fill_img[~nz]


# ----------------------------------------------------------------------------
# orchestration  (lib/vnlb/search/search.py, proc_nl.py, impl.py)
# ----------------------------------------------------------------------------

def _torch_randperm(n):
    import torch
    return torch.randperm(n).numpy()


def proc_nl(noisy, basic, flows, args, randperm=_torch_randperm, stats=None,
            dist_chnls=None, window_mode="shift"):
    """One VNLB step: lib/vnlb/proc_nl.py:38-141 with exec_search
    (lib/vnlb/search/search.py:25-69) inlined.  noisy/basic RGB [T,C,H,W] f32;
    returns the step's RGB output."""
    shape = noisy.shape
    T, C, H, W = shape
    mask, _ = init_mask(shape, args)                                   # :44
    tsize, k = args.tsize, args.npatches
    p_noisy = np.zeros((tsize, k, args.pt, C, args.ps, args.ps), np.float32)   # alloc.py:10-30
    p_basic = np.zeros_like(p_noisy)
    flat = np.zeros(tsize, bool)
    vals = np.zeros((tsize, k), np.float32)                           # alloc.py:74-87
    inds = -np.ones((tsize, k), np.int64)
    nelems = int(mask.sum())
    nbatches = (nelems - 1) // tsize + 1                              # batching.py:10-16
    y_noisy = rgb2yuv(noisy)                                          # :57
    y_basic = rgb2yuv(basic)
    deno = np.zeros(shape, np.float32)                                # rgb2yuv(0) == 0
    weights = np.zeros((T, H, W), np.float32)
    srch = y_noisy if args.srch_img == "noisy" else y_basic
    ngroups = 0
    for _ in range(nbatches):                                         # :64
        # ---- exec_search, search.py:25-69 ----
        inds[...] = -1
        vals[...] = np.inf
        done = False
        for index in range(args.nstreams):
            q = mask2inds(mask, args.bsize, randperm)
            if q.shape[0] == 0:
                done = True
                break
            sl = slice(index * args.bsize, (index + 1) * args.bsize)  # batching.py:25-28
            exec_sim_search_burst(srch, q, vals[sl], inds[sl], flows, args.sigma, args,
                                  dist_chnls, window_mode)
            fill_patches(p_noisy[sl], y_noisy, inds[sl])
            fill_patches(p_basic[sl], y_basic, inds[sl])
            update_mask_inds(mask, inds[sl], C)
        done = done or mask.sum() == 0
        # ---- proc_nl.py:73-87 ----
        if args.step == 1:
            flat[...] = exec_flat_areas(p_noisy, args.gamma, args.sigma2)
        valid = np.all(inds != -1, 1)
        if valid.sum() == 0:
            break
        ngroups += int(valid.sum())
        out_n, out_b, _ = bayes_denoise(p_noisy[valid], p_basic[valid], flat[valid], args)
        p_noisy[valid] = out_n
        p_basic[valid] = out_b
        agg_patches(deno, weights, p_noisy, inds)
        if done:
            break
    normalize(deno, weights, y_basic if args.step == 1 else y_noisy)   # :118-125
    if stats is not None:
        stats.setdefault("ngroups", []).append(ngroups)
    return yuv2rgb(deno)                                              # :138


def denoise(noisy, sigma, flows=None, params=None, stats=None, dist_chnls=None,
            window_mode="shift", randperm=_torch_randperm):
    """vnlb.denoise, lib/vnlb/impl.py:24-62 with `default_params` (classic
    VNLB settings).  Returns (deno, basic, seconds)."""
    t0 = time.time()
    noisy = np.ascontiguousarray(noisy, dtype=np.float32)
    c = noisy.shape[1]
    params = params or default_params(sigma)
    basic = proc_nl(noisy, np.zeros_like(noisy), flows, get_args(params, c, 0),
                    randperm, stats, dist_chnls, window_mode)
    deno = proc_nl(noisy, basic, flows, get_args(params, c, 1),
                   randperm, stats, dist_chnls, window_mode)
    return deno, basic, time.time() - t0


def compute_psnrs(deno, clean, imax=255.):
    """lib/vnlb/utils/metrics.py:50-71."""
    d = np.asarray(deno, np.float64) / imax
    c = np.asarray(clean, np.float64) / imax
    return -10 * np.log10(((d - c) ** 2).mean(axis=(-3, -2, -1)))


# ----------------------------------------------------------------------------
# synthetic data: shared by tests and bench
# ----------------------------------------------------------------------------

def synth_video(T, H, W, seed=123, C=3, return_flows=False, crop=None):
    """Deterministic clean video: smooth field + textured rectangles that
    translate 1-2 px/frame; range 0..255, float32 [T,C,H,W].  Also returns the
    analytic forward/backward flows [T,2,H,W] (ch0 = dx, ch1 = dy) of the
    background (zero) -- per-object motion is small and only approximate."""
    rng = np.random.RandomState(seed)
    # crop = (Hc, Wc): only the top-left Hc x Wc corner of the H x W video is materialised (same content as cropping
    # the full video; the CPU baseline's bounded sample of a 1080p workload)
    Hc, Wc = (H, W) if crop is None else (min(int(crop[0]), H), min(int(crop[1]), W))
    yy, xx = np.meshgrid(np.arange(Hc, dtype=np.float32), np.arange(Wc, dtype=np.float32), indexing="ij")
    vid = np.zeros((T, C, Hc, Wc), np.float32)
    ff = np.zeros((T, 2, Hc, Wc), np.float32) if return_flows else None
    bf = np.zeros((T, 2, Hc, Wc), np.float32) if return_flows else None
    nrect = 6
    rects = []
    for _ in range(nrect):
        rh, rw = rng.randint(H // 6 + 2, H // 2 + 3), rng.randint(W // 6 + 2, W // 2 + 3)
        y0, x0 = rng.randint(0, max(1, H - rh)), rng.randint(0, max(1, W - rw))
        vy, vx = rng.randint(-2, 3), rng.randint(-2, 3)
        tex = rng.rand(C, rh, rw).astype(np.float32) * 60 + rng.rand(C, 1, 1).astype(np.float32) * 150
        fy, fx = rng.uniform(0.2, 1.2), rng.uniform(0.2, 1.2)
        stripes = 25 * np.sin(fy * np.arange(rh)[:, None] + fx * np.arange(rw)[None, :]).astype(np.float32)
        rects.append((y0, x0, rh, rw, vy, vx, tex * 0.3 + stripes[None] + 60))
    for t in range(T):
        for ch in range(C):
            vid[t, ch] = 110 + 60 * np.sin(xx / (17. + 3 * ch) + 0.1 * t) * np.cos(yy / (23. - 2 * ch))
        for (y0, x0, rh, rw, vy, vx, tex) in rects:
            ya, xa = y0 + vy * t, x0 + vx * t
            ys, xs = max(0, ya), max(0, xa)
            ye, xe = min(Hc, ya + rh), min(Wc, xa + rw)
            if ye > ys and xe > xs:
                vid[t, :, ys:ye, xs:xe] = tex[:, ys - ya:ye - ya, xs - xa:xe - xa]
                if return_flows:     # the object's own translation (the last-drawn object wins, as in the frame)
                    ff[t, 0, ys:ye, xs:xe], ff[t, 1, ys:ye, xs:xe] = vx, vy
                    bf[t, 0, ys:ye, xs:xe], bf[t, 1, ys:ye, xs:xe] = -vx, -vy
    vid = np.clip(vid, 0, 255).astype(np.float32)
    if return_flows:     # "precomputed flows" of SURVEY 8d: the generator's analytic translation field, |flow| <= 2 px/frame
        return vid, dict(fflow=ff, bflow=bf)
    return vid


def add_noise(clean, sigma, seed=123):
    rng = np.random.RandomState(seed + 1)
    return (clean + rng.randn(*clean.shape).astype(np.float32) * np.float32(sigma)).astype(np.float32)
