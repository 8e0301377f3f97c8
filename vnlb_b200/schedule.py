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
            if len(cls._cache) >= 4:          # both steps (k = 100 / 60) keep their buffers between calls
                cls._cache.clear()
            ws = AttrDict()
            rows = cap if stacks else 0     # the fused kernel never materialises the patch stacks
            ws.noisy = torch.empty((rows, k, pt, c, ps, ps), dtype=torch.float32, device=device)
            ws.basic = torch.empty((rows, k, pt, c, ps, ps), dtype=torch.float32, device=device)
            ws.flat = torch.zeros((cap,), dtype=torch.uint8, device=device)
            # two sets of round buffers: the search of round r+1 overlaps the Bayes kernel of round r
            ws.vals2 = [torch.empty((cap, k), dtype=torch.float32, device=device) for _ in range(2)]
            ws.inds2 = [torch.empty((cap, k), dtype=torch.int64, device=device) for _ in range(2)]
            ws.qinds2 = [torch.empty((cap, 3), dtype=torch.int64, device=device) for _ in range(2)]
            ws.vals, ws.inds, ws.qinds = ws.vals2[0], ws.inds2[0], ws.qinds2[0]
            ws.counters4 = torch.zeros((4, 2), dtype=torch.int32, device=device)
            ws.host4 = torch.zeros((4, 2), dtype=torch.int32).pin_memory()
            ws.ev_search = [torch.cuda.Event() for _ in range(2)]     # re-used every round (no per-round garbage)
            ws.ev_bayes = [torch.cuda.Event() for _ in range(2)]
            ws.ev_copied = [torch.cuda.Event() for _ in range(4)]
            ws.search_stream = torch.cuda.Stream(device=device)
            ws.bayes_stream = torch.cuda.Stream(device=device)
            ws.counters = torch.zeros((2,), dtype=torch.int32, device=device)
            ws.host = torch.zeros((2,), dtype=torch.int32).pin_memory()
            cls._cache[key] = ws
        return ws


def _rounds_overlapped(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed):
    """The rounds of one step on two streams: [count, draw, search, clear mask] of round r+1 runs on
    the search stream while the fused Bayes kernel of round r runs on the Bayes stream (the mask only
    depends on the search results, never on the filtered patches)."""
    t, c, h, w = images.shape
    main = torch.cuda.current_stream()
    sA, sB = ws.search_stream, ws.bayes_stream
    sA.wait_stream(main)
    sB.wait_stream(main)
    done_search = [None, None]      # event: inds2[b] written and mask updated
    done_bayes = [None, None]       # event: inds2[b] consumed
    nproc, nrounds, nmask0, buf = 0, 0, None, 0
    tm = L.timer
    while True:
        with torch.cuda.stream(sA):
            st = L.stream_ptr()
            ws.counters.zero_()
            L.check(L.lib.vnlb_count_mask(L.ptr(mask), t, h, w, L.ptr(ws.counters), st), "vnlb_count_mask")
            ws.host.copy_(ws.counters, non_blocking=True)
            sA.synchronize()
            remaining = int(ws.host[0])
            if nmask0 is None:
                nmask0 = remaining
                qmin = min(qmin, max(296, nmask0 // 64))     # small videos: small rounds keep the greedy mask effective
            if remaining == 0:
                break
            target = min(cap, max(qmin, int(remaining * frac)))
            prob = 1.0 if remaining <= target else target / remaining * 0.97   # stay under cap
            if done_bayes[buf] is not None:
                sA.wait_event(done_bayes[buf])                               # round r-2 is done with this buffer
            qinds, vals, inds = ws.qinds2[buf], ws.vals2[buf], ws.inds2[buf]
            L.check(L.lib.vnlb_select_queries(L.ptr(mask), t, h, w, prob, seed, nrounds, L.ptr(qinds), cap,
                                              L.ptr(ws.counters), st), "vnlb_select_queries")
            ws.host.copy_(ws.counters, non_blocking=True)
            sA.synchronize()
            q = min(int(ws.host[1]), cap)
            nrounds += 1
            if q == 0:
                continue
            tok = tm.start("search_s%d" % args.step) if tm else None
            search.exec_sim_search_burst(srch_img, qinds[:q], vals[:q], inds[:q], flows, args.sigma, args)
            if tm:
                tm.stop(tok)
            search_mask.update_mask_inds(mask, inds[:q], c, boost=args.aggreBoost)
            done_search[buf] = torch.cuda.Event()
            done_search[buf].record(sA)
        with torch.cuda.stream(sB):
            sB.wait_event(done_search[buf])
            tok = tm.start("bayes_s%d" % args.step) if tm else None
            deno.bayes_aggregate_fused(images, inds[:q], args)
            if tm:
                tm.stop(tok)
            done_bayes[buf] = torch.cuda.Event()
            done_bayes[buf].record(sB)
        nproc += q
        buf ^= 1
    main.wait_stream(sA)
    main.wait_stream(sB)
    return nproc, nrounds, nmask0


def _rounds_async(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed, est_mask, row_hist=None):
    """Like _rounds_overlapped, but the host never waits for the GPU inside the loop: every kernel of a
    round is enqueued with `cap` rows (rows beyond the number actually drawn are padded to invalid queries
    and skipped on the device), the round size is read back two rounds late from a ring of pinned counters,
    and the draw probability is computed from that (stale, hence conservative) count.  The loop ends when a
    read-back says the mask is empty; the one or two extra rounds already enqueued are no-ops."""
    t, c, h, w = images.shape
    main = torch.cuda.current_stream()
    sA, sB = ws.search_stream, ws.bayes_stream
    sA.wait_stream(main)
    sB.wait_stream(main)
    done_search, done_bayes, copied = ws.ev_search, ws.ev_bayes, ws.ev_copied
    used_bayes = [False, False]
    nproc, nrounds, nmask0 = 0, 0, None
    remaining = int(est_mask)            # until the first read-back: analytic size of the lattice
    qmin = min(qmin, max(296, remaining // 64))
    tm = L.timer
    rows_of = [cap] * 4
    r = 0
    while True:
        buf, slot = r & 1, r & 3
        target = min(cap, max(qmin, int(remaining * frac)))
        prob = 1.0 if remaining <= target else target / remaining * 0.97
        # rows enqueued this round: the expected draw is <= 0.97 target (the count is stale, hence an over-estimate), so
        # target + 25 % bounds it; pixels the selection cannot place stay masked for a later round.  Late rounds thus
        # launch a third of the CTAs of a `cap`-sized round instead of thousands that exit at once.
        rows = min(cap, int(target * 1.25) + 256)
        rows_of[slot] = rows
        with torch.cuda.stream(sA):
            st = L.stream_ptr()
            cnt = ws.counters4[slot]
            cnt.zero_()
            if used_bayes[buf]:
                sA.wait_event(done_bayes[buf])                               # round r-2 is done with this buffer
            qinds, vals, inds = ws.qinds2[buf][:rows], ws.vals2[buf][:rows], ws.inds2[buf][:rows]
            L.check(L.lib.vnlb_count_mask(L.ptr(mask), t, h, w, L.ptr(cnt), st), "vnlb_count_mask")
            L.check(L.lib.vnlb_select_queries(L.ptr(mask), t, h, w, prob, seed, r, L.ptr(qinds), rows, L.ptr(cnt), st),
                    "vnlb_select_queries")
            L.check(L.lib.vnlb_pad_queries(L.ptr(qinds), L.ptr(cnt), rows, st), "vnlb_pad_queries")
            if row_hist is not None:     # groups per image row (multi-GPU: balances the bands of the next step)
                yq = qinds[:, 1]
                row_hist.index_add_(0, yq.clamp(min=0), (yq >= 0).to(row_hist.dtype))
            ws.host4[slot].copy_(cnt, non_blocking=True)
            copied[slot].record(sA)
            tok = tm.start("search_s%d" % args.step) if tm else None
            search.exec_sim_search_burst(srch_img, qinds, vals, inds, flows, args.sigma, args)
            if tm:
                tm.stop(tok)
            search_mask.update_mask_inds(mask, inds, c, boost=args.aggreBoost)
            done_search[buf].record(sA)
        with torch.cuda.stream(sB):
            sB.wait_event(done_search[buf])
            tok = tm.start("bayes_s%d" % args.step) if tm else None
            deno.bayes_aggregate_fused(images, inds, args)
            if tm:
                tm.stop(tok)
            done_bayes[buf].record(sB)
            used_bayes[buf] = True
        r += 1
        if r >= 2:                       # read back round r-2 (long finished: no stall in steady state)
            old = (r - 2) & 3
            copied[old].synchronize()
            rem_old, nsel = int(ws.host4[old][0]), min(int(ws.host4[old][1]), rows_of[old])
            if nmask0 is None:
                nmask0 = rem_old
            nproc += nsel
            nrounds += 1
            if rem_old == 0:
                break
            remaining = max(rem_old - nsel, 1)
    # the last enqueued round (r-1) drew nothing or its groups are counted here
    last = (r - 1) & 3
    copied[last].synchronize()
    nproc += min(int(ws.host4[last][1]), rows_of[last])
    main.wait_stream(sA)
    main.wait_stream(sB)
    return nproc, nrounds, nmask0


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
    mask = search_mask.init_mask_device(images.shape, args, dev, y_range)
    st = L.stream_ptr()
    color.rgb2yuv_images(images)
    srch_img = {"noisy": images.noisy, "basic": images.basic, "clean": images.clean}[args.srch_img]
    if srch_img is None:
        raise ValueError("uknown search image [%s]" % args.srch_img)
    nproc, nrounds, nmask0 = 0, 0, None
    overlap = args.get("fast_overlap", "async")
    if fused and overlap:
        if overlap == "async":
            bands = [(0, h)] if y_range is None else ([y_range] if isinstance(y_range[0], int) else list(y_range))
            rows = sum(min(b, h - args.ps + 1) - a for a, b in bands)
            est = (t - args.pt + 1) * max(rows, 1) * (w - args.ps + 1) // (args.procStep ** 2)
            row_hist = None
            if stats is not None and stats.get("want_row_hist"):
                row_hist = torch.zeros((h,), dtype=torch.float32, device=dev)
                stats["row_hist"] = row_hist
            nproc, nrounds, nmask0 = _rounds_async(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed,
                                                   est + est // 8, row_hist)
        else:
            nproc, nrounds, nmask0 = _rounds_overlapped(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed)
        finish_step(images, args, reduce_fn)
        if stats is not None:
            stats.setdefault("ngroups", []).append(nproc)
            stats.setdefault("nmask", []).append(nmask0)
            stats.setdefault("nrounds", []).append(nrounds)
        return
    while True:
        ws.counters.zero_()
        L.check(L.lib.vnlb_count_mask(L.ptr(mask), t, h, w, L.ptr(ws.counters), st), "vnlb_count_mask")
        ws.host.copy_(ws.counters, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        remaining = int(ws.host[0])
        if nmask0 is None:
            nmask0 = remaining
            qmin = min(qmin, max(296, nmask0 // 64))
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
        tok = tm.start("search_s%d" % args.step) if tm else None
        search.exec_sim_search_burst(srch_img, ws.qinds[:q], vals, inds, flows, args.sigma, args)
        if tm:
            tm.stop(tok)
            tok = tm.start("mask_fill_flat")
        search_mask.update_mask_inds(mask, inds, c, boost=args.aggreBoost)
        if fused:
            if tm:
                tm.stop(tok)
                tok = tm.start("bayes_s%d" % args.step)
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
            tok = tm.start("bayes_s%d" % args.step)
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
