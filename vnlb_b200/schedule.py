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
one sub-batch of 128): vnlb_round_dedup drops the later one before it is
filtered, which keeps the number of groups at the reference schedule's level.
One 8-byte device->host read per round, consumed two rounds late, tells the host
the round size; the host never waits for the GPU inside the loop."""
import threading
from collections import OrderedDict

import torch

from . import _lib as L
from . import agg, color, deno, search
from . import mask as search_mask
from .flat_areas import update_flat_patch
from .proc_nl import finish_step
from .utils import AttrDict

FAST_DEFAULTS = dict(fast_frac=1. / 8, fast_min=4096, fast_cap=16384, fast_seed=123, fast_dedup=True)


class _Workspace:
    """Round buffers, streams and events, cached per (device, rows, k, pt, c, ps) so both steps and repeated calls
    reuse them (LRU, at most `_max` entries; guarded by a lock).  denoise() is NOT re-entrant per device: two host
    threads running the fast schedule on the same device at the same time would share these buffers."""
    _cache = OrderedDict()
    _lock = threading.Lock()
    _max = 8

    @classmethod
    def get(cls, device, cap, k, pt, c, ps, stacks=True):
        key = (str(device), cap, k, pt, c, ps, stacks)
        with cls._lock:
            ws = cls._cache.get(key)
            if ws is not None:
                cls._cache.move_to_end(key)
                return ws
            while len(cls._cache) >= cls._max:
                cls._cache.popitem(last=False)
            ws = cls._cache[key] = cls._make(device, cap, k, pt, c, ps, stacks)
            return ws

    @staticmethod
    def _make(device, cap, k, pt, c, ps, stacks):
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
        ws.graph_stream = torch.cuda.Stream(device=device)      # fast_graph: capture + replay stream
        ws.round_state = torch.zeros((4,), dtype=torch.int32, device=device)   # device-side round index (vnlb_round_draw)
        ws.dropped = torch.zeros((1,), dtype=torch.int32, device=device)
        ws.counters = torch.zeros((2,), dtype=torch.int32, device=device)
        ws.host = torch.zeros((2,), dtype=torch.int32).pin_memory()
        return ws


def _rounds_async(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed, est_mask, row_hist=None, sync_hook=None):
    """The rounds of one step on two streams: [count, draw, search, clear mask] of round r+1 runs on the search
    stream while the fused Bayes kernels of round r run on the Bayes stream (the mask only depends on the search
    results, never on the filtered patches).  The host never waits for the GPU inside the loop: every kernel of a
    round is enqueued with `cap` rows (rows beyond the number actually drawn are padded to invalid queries
    and skipped on the device), the round size is read back two rounds late from a ring of pinned counters,
    and the draw probability is computed from that (stale, hence conservative) count.  The loop ends when a
    read-back says the mask is empty; the one or two extra rounds already enqueued are no-ops.
    `sync_hook` = (round index R, fn): multi-GPU load balancing -- before round R is enqueued (or after the loop, if it
    ends earlier) the streams are drained and fn(mask) runs exactly once; it may rewrite rows of the mask."""
    t, c, h, w = images.shape
    main = torch.cuda.current_stream()
    sA, sB = ws.search_stream, ws.bayes_stream
    # greedy conflict resolution inside a round (vnlb_round_dedup): a drawn pixel that an earlier row of the same round
    # covers is not processed, as in the reference's sequential sub-batches
    dedup = bool(args.get("fast_dedup", FAST_DEFAULTS["fast_dedup"]))
    owner = torch.full((t, h, w), -1, dtype=torch.int32, device=images.device) if dedup else None
    ws.dropped.zero_()
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
    hook_done = sync_hook is None
    while True:
        if not hook_done and r == sync_hook[0]:
            hook_done = True
            sA.synchronize()
            sB.synchronize()
            with torch.cuda.stream(sA):
                sync_hook[1](mask)
                remaining = max(int(mask.sum().item()), 1)
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
            if dedup:
                L.check(L.lib.vnlb_round_dedup(L.ptr(qinds), L.ptr(inds), rows, int(inds.shape[1]), L.ptr(owner), r,
                                               L.ptr(mask), t, c, h, w, int(bool(args.aggreBoost)), L.ptr(ws.dropped), st),
                        "vnlb_round_dedup")
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
    if not hook_done:                    # the step ended before round R: still take part in the collective, then finish
        sA.synchronize()
        sB.synchronize()
        with torch.cuda.stream(sA):
            sync_hook[1](mask)
            left = int(mask.sum().item())
        if left > 0:                     # rows received from a neighbour: process them with the plain loop
            ndrop1 = int(ws.dropped.item()) if dedup else 0
            main.wait_stream(sA)
            main.wait_stream(sB)
            n2, r2, _, d2 = _rounds_async(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed + 1, left, row_hist, None)
            return nproc - ndrop1 + n2, nrounds + r2, nmask0, ndrop1 + d2
    main.wait_stream(sA)
    main.wait_stream(sB)
    ndrop = int(ws.dropped.item()) if dedup else 0          # (end of the step: the one host sync of the loop)
    return nproc - ndrop, nrounds, nmask0, ndrop


def _graph_rows(rows_needed, cap):
    """Rows of a captured round: the smallest of cap, 3/4 cap, 1/2 cap, 3/8 cap, ... (>= 512) that holds the draw."""
    r = cap
    while True:
        for cand in (r * 3 // 4, r // 2):
            if cand < 512 or cand < rows_needed:
                return r
            r = cand


def _rounds_graph(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed, est_mask):
    """The rounds of one step as CUDA-graph replays (args.fast_graph).  A round is a device-controlled program --
    vnlb_round_draw (count, draw with the probability computed on the device from the live count, pad), search,
    in-round conflict resolution, mask update, fused Bayes + aggregation -- whose launches depend on no host scalar
    but the number of rows, so it is captured once per row bucket (_graph_rows: at most ~10 graphs per step) and
    replayed for every round that fits the bucket: two enqueues per round (the replay and the 8-byte read-back)
    instead of fifteen launches.  The first round runs eagerly (it allocates what the wrappers allocate lazily: graph
    capture must not allocate).  The host reads the round size two rounds late, as in _rounds_async, to size the next
    bucket and to see the mask run empty.  Search and Bayes of consecutive rounds do not overlap in this mode.  The
    draw uses the live count where the host loop uses a count that is two rounds old: same rule, slightly different
    draws (both deterministic)."""
    t, c, h, w = images.shape
    main = torch.cuda.current_stream()
    sG = ws.graph_stream
    dedup = bool(args.get("fast_dedup", FAST_DEFAULTS["fast_dedup"]))
    owner = torch.full((t, h, w), -1, dtype=torch.int32, device=images.device) if dedup else None
    ws.dropped.zero_()
    ws.round_state.zero_()
    sG.wait_stream(main)
    copied = ws.ev_copied
    cnt = ws.counters4[0]
    qbuf, vbuf, ibuf = ws.qinds2[0], ws.vals2[0], ws.inds2[0]
    k = int(ibuf.shape[1])
    remaining = int(est_mask)
    qmin = min(qmin, max(296, remaining // 64))
    graphs = {}
    rows_of = [cap] * 4
    nproc, nrounds, nmask0 = 0, 0, None

    def enqueue_round(rows):
        st = L.stream_ptr()
        qinds, vals, inds = qbuf[:rows], vbuf[:rows], ibuf[:rows]
        L.check(L.lib.vnlb_round_draw(L.ptr(mask), t, h, w, float(frac), int(qmin), int(rows), seed, L.ptr(ws.round_state),
                                      L.ptr(qinds), L.ptr(cnt), st), "vnlb_round_draw")
        search.exec_sim_search_burst(srch_img, qinds, vals, inds, flows, args.sigma, args)
        if dedup:
            L.check(L.lib.vnlb_round_dedup_dev(L.ptr(qinds), L.ptr(inds), rows, k, L.ptr(owner), L.ptr(ws.round_state),
                                               L.ptr(mask), t, c, h, w, int(bool(args.aggreBoost)), L.ptr(ws.dropped), st),
                    "vnlb_round_dedup_dev")
        search_mask.update_mask_inds(mask, inds, c, boost=args.aggreBoost)
        deno.bayes_aggregate_fused(images, inds, args)

    r = 0
    with torch.cuda.stream(sG):
        while True:
            slot = r & 3
            target = min(cap, max(qmin, int(remaining * frac)))
            rows = _graph_rows(min(cap, int(target * 1.25) + 256), cap)
            rows_of[slot] = rows
            if r == 0:
                enqueue_round(rows)
            else:
                g = graphs.get(rows)
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    g.capture_begin()
                    try:
                        enqueue_round(rows)
                    finally:
                        g.capture_end()
                    graphs[rows] = g
                g.replay()
            ws.host4[slot].copy_(cnt, non_blocking=True)
            copied[slot].record(sG)
            r += 1
            if r >= 2:
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
        last = (r - 1) & 3
        copied[last].synchronize()
        nproc += min(int(ws.host4[last][1]), rows_of[last])
    main.wait_stream(sG)
    ndrop = int(ws.dropped.item()) if dedup else 0
    del graphs
    return nproc - ndrop, nrounds, nmask0, ndrop


def proc_nl_fast(images, flows, args, stats=None, y_range=None, reduce_fn=None, tile=None, post_fn=None, sync_hook=None):
    """One VNLB step with the throughput schedule (same contract as proc_nl).  `y_range` / `tile`: multi-GPU band of
    reference rows and position of this row tile in the frame (mask.init_mask_device)."""
    dev = images.device
    t, c, h, w = images.shape
    frac = float(args.get("fast_frac", FAST_DEFAULTS["fast_frac"]))
    qmin = int(args.get("fast_min", FAST_DEFAULTS["fast_min"]))
    cap = int(args.get("fast_cap", FAST_DEFAULTS["fast_cap"]))
    seed = int(args.get("fast_seed", FAST_DEFAULTS["fast_seed"])) + 7919 * int(args.step)
    k = args.npatches
    # the fused kernel keeps 32-bit image offsets and needs its patch shape; anything else runs the staged operators
    fused = bool(args.get("fused", True)) and deno.fused_supported(args, c, images.shape)
    ws = _Workspace.get(dev, cap, k, args.pt, c, args.ps, stacks=not fused)
    mask = search_mask.init_mask_device(images.shape, args, dev, y_range, tile)
    st = L.stream_ptr()
    if not images.get("is_yuv"):
        color.rgb2yuv_images(images)
    srch_img = {"noisy": images.noisy, "basic": images.basic, "clean": images.clean}[args.srch_img]
    if srch_img is None:
        raise ValueError("uknown search image [%s]" % args.srch_img)
    nproc, nrounds, nmask0 = 0, 0, None
    if fused and args.get("fast_overlap", "async"):
        bands = [(0, h)] if y_range is None else ([y_range] if isinstance(y_range[0], int) else list(y_range))
        rows = sum(min(b, h - args.ps + 1) - a for a, b in bands)
        est = (t - args.pt + 1) * max(rows, 1) * (w - args.ps + 1) // (args.procStep ** 2)
        row_hist = None
        if stats is not None and stats.get("want_row_hist"):
            row_hist = torch.zeros((h,), dtype=torch.float32, device=dev)
            stats["row_hist"] = row_hist
        if args.get("fast_graph", False) and sync_hook is None and row_hist is None and L.timer is None:
            nproc, nrounds, nmask0, ndrop = _rounds_graph(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed,
                                                          est + est // 8)
        else:
            nproc, nrounds, nmask0, ndrop = _rounds_async(images, flows, args, ws, mask, srch_img, frac, qmin, cap, seed,
                                                          est + est // 8, row_hist, sync_hook)
        finish_step(images, args, reduce_fn, post_fn)
        if stats is not None:
            stats.setdefault("ndropped", []).append(ndrop)
            stats.setdefault("ngroups", []).append(nproc)
            stats.setdefault("nmask", []).append(nmask0)
            stats.setdefault("nrounds", []).append(nrounds)
        return
    if images.basic is None:            # staged operators gather the (all-zero) basic image in step 1 like the reference
        images.basic = torch.zeros_like(images.noisy)
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
    finish_step(images, args, reduce_fn, post_fn)
    if stats is not None:
        stats.setdefault("ngroups", []).append(nproc)
        stats.setdefault("nmask", []).append(nmask0)
        stats.setdefault("nrounds", []).append(nrounds)
