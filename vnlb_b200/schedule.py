"""Throughput ("fast") schedule of one VNLB step.

Same algorithm, parameters and kernels as the parity schedule (proc_nl.py); what
changes is WHICH masked pixels become reference pixels and how many are handled
per launch.  The reference draws 128 pixels per sub-batch with a host randperm
and re-reads the mask between sub-batches (lib/vnlb/search/search.py:38-64,
lib/vnlb/search_mask/mask.py:18-31): 3+ host syncs and a full-mask compaction
per 128 groups.  Here each round draws, on the device, a hash-random fraction
`fast_frac` of the remaining masked pixels (at least `fast_min`, at most
`fast_cap` rows), searches them in one launch, clears everything they found
(paste trick + boost) and then filters/aggregates the round.  Two pixels drawn
in the same round may cover each other (the reference has the same effect inside
one sub-batch); a small `fast_frac` keeps that redundancy at a few percent.
One 8-byte device->host read per round tells the host the round size."""
import torch

from . import _lib as L
from . import agg, color, deno, search
from . import mask as search_mask
from .flat_areas import update_flat_patch
from .proc_nl import finish_step
from .utils import AttrDict

FAST_DEFAULTS = dict(fast_frac=1. / 8, fast_min=4096, fast_cap=16384, fast_seed=123)


class _Workspace:
    """Round buffers, cached per (device, rows, k, pt, c, ps) so both steps and
    repeated calls reuse them."""
    _cache = {}

    @classmethod
    def get(cls, device, cap, k, pt, c, ps, stacks=True):
        key = (str(device), cap, k, pt, c, ps, stacks)
        ws = cls._cache.get(key)
        if ws is None:
            cls._cache.clear()
            ws = AttrDict()
            rows = cap if stacks else 0     # the fused kernel never materialises the patch stacks
            ws.noisy = torch.empty((rows, k, pt, c, ps, ps), dtype=torch.float32, device=device)
            ws.basic = torch.empty((rows, k, pt, c, ps, ps), dtype=torch.float32, device=device)
            ws.flat = torch.zeros((cap,), dtype=torch.uint8, device=device)
            ws.vals = torch.empty((cap, k), dtype=torch.float32, device=device)
            ws.inds = torch.empty((cap, k), dtype=torch.int64, device=device)
            ws.qinds = torch.empty((cap, 3), dtype=torch.int64, device=device)
            ws.counters = torch.zeros((2,), dtype=torch.int32, device=device)
            ws.host = torch.zeros((2,), dtype=torch.int32).pin_memory()
            cls._cache[key] = ws
        return ws


def proc_nl_fast(images, flows, args, stats=None, y_range=None, reduce_fn=None):
    """One VNLB step with the throughput schedule (same contract as proc_nl)."""
    dev = images.device
    t, c, h, w = images.shape
    frac = float(args.get("fast_frac", FAST_DEFAULTS["fast_frac"]))
    qmin = int(args.get("fast_min", FAST_DEFAULTS["fast_min"]))
    cap = int(args.get("fast_cap", FAST_DEFAULTS["fast_cap"]))
    seed = int(args.get("fast_seed", FAST_DEFAULTS["fast_seed"])) + 7919 * int(args.step)
    k = args.npatches
    fused = bool(args.get("fused", True)) and deno.fused_supported(args, c)
    ws = _Workspace.get(dev, cap, k, args.pt, c, args.ps, stacks=not fused)
    mask = torch.empty((t, h, w), dtype=torch.int8, device=dev)
    y0, y1 = (0, h) if y_range is None else y_range
    st = L.stream_ptr()
    L.check(L.lib.vnlb_init_mask(L.ptr(mask), t, h, w, args.ps, args.pt, args.procStep, int(y0), int(y1), st),
            "vnlb_init_mask")
    color.rgb2yuv_images(images)
    srch_img = {"noisy": images.noisy, "basic": images.basic, "clean": images.clean}[args.srch_img]
    if srch_img is None:
        raise ValueError("uknown search image [%s]" % args.srch_img)
    nproc, nrounds, nmask0 = 0, 0, None
    while True:
        ws.counters.zero_()
        L.check(L.lib.vnlb_count_mask(L.ptr(mask), t, h, w, L.ptr(ws.counters), st), "vnlb_count_mask")
        ws.host.copy_(ws.counters, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        remaining = int(ws.host[0])
        if nmask0 is None:
            nmask0 = remaining
        if remaining == 0:
            break
        target = min(cap, max(qmin, int(remaining * frac)))
        prob = 1.0 if remaining <= target else target / remaining * 0.97   # stay under cap
        L.check(L.lib.vnlb_select_queries(L.ptr(mask), t, h, w, prob, seed, nrounds, L.ptr(ws.qinds), cap,
                                          L.ptr(ws.counters), st), "vnlb_select_queries")
        ws.host.copy_(ws.counters, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        q = min(int(ws.host[1]), cap)
        nrounds += 1
        if q == 0:
            continue
        rows = AttrDict(noisy=ws.noisy[:q], basic=ws.basic[:q], flat=ws.flat[:q], clean=None)
        vals, inds = ws.vals[:q], ws.inds[:q]
        tm = L.timer
        tok = tm.start("search") if tm else None
        search.exec_sim_search_burst(srch_img, ws.qinds[:q], vals, inds, flows, args.sigma, args)
        if tm:
            tm.stop(tok)
            tok = tm.start("mask_fill_flat")
        search_mask.update_mask_inds(mask, inds, c, boost=args.aggreBoost)
        if fused:
            if tm:
                tm.stop(tok)
                tok = tm.start("bayes")
            deno.bayes_aggregate_fused(images, inds, args)
            if tm:
                tm.stop(tok)
            nproc += q
            continue
        search.fill_patches(rows.noisy, images.noisy, inds)
        if args.step == 1:
            search.fill_patches(rows.basic, images.basic, inds)
        update_flat_patch(rows, args, inds)
        if tm:
            tm.stop(tok)
            tok = tm.start("bayes")
        deno.denoise(rows, args, args.deno, inds)
        if tm:
            tm.stop(tok)
            tok = tm.start("aggregate")
        agg.compute_agg_batch(images.deno, rows.noisy, inds, images.weights, None, None, args.ps, args.pt)
        if tm:
            tm.stop(tok)
        nproc += q
    finish_step(images, args, reduce_fn)
    if stats is not None:
        stats.setdefault("ngroups", []).append(nproc)
        stats.setdefault("nmask", []).append(nmask0)
        stats.setdefault("nrounds", []).append(nrounds)
