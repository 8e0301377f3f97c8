"""Multi-GPU VNLB: one process per GPU (torch.distributed, NCCL over NVLink), the video split into row bands with halos.

The path shards over REFERENCE PIXELS (BASELINE north star): rank r owns one contiguous band
[y0_r, y1_r) of reference rows and holds only the row TILE [y0_r - halo, y1_r + halo) of the video (and of the
flows) in HBM -- the rows its search windows and patches can reach:

    halo = w_s // 2 + ps - 1 + ceil(nWt * (max|flow_y| + 0.5))      (19 rows for the default parameters, zero flow)

so each rank copies (band + 2 halo) / H of the video host -> device, runs the single-GPU step on its tile (the same
kernels, tile-local coordinates, the reference-pixel lattice kept in GLOBAL phase by vnlb_init_mask_tile) and
exchanges, per step, only what overlaps a neighbour:

  1. the accumulator rows (sum image + weights) its groups wrote into rows another rank owns are sent to that
     rank and added there (neighbour send/recv of T x (C+1) x rows x W floats per border and direction);
  2. after normalisation the owner returns the normalised rows of step 1 that lie in the neighbour's halo, because
     step 2 searches and gathers the `basic` estimate over the whole tile.

No other data-path communication: there is no all-reduce of frames.  At the end the bands are gathered so that every
rank returns the full (deno, basic) like the single-GPU call.  Each rank runs its own greedy mask over its own band,
so a group found across a band border does not clear the other rank's mask: slightly more groups near the N-1
borders and an estimate there that differs from the single-GPU one within the PSNR tolerance (0.02 dB), not
max-abs (caveat H7 of SURVEY.md; tests/test_dist_gpu.py asserts it)."""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import alloc, color
from .params import get_args, get_params
from .utils import AttrDict, Timer, expand_flows


# ------------------------------------------------------------------------------------------------ partition (host)
def partition_rows(h, ps, world_size, rank):
    """Band [y0, y1) of reference-pixel rows (valid rows are 0..h-ps) of `rank`: equal row counts."""
    valid = h - ps + 1
    y0 = (valid * rank) // world_size
    y1 = (valid * (rank + 1)) // world_size
    if rank == world_size - 1:
        y1 = h
    return y0, y1


def partition_rows_weighted(weights, ps, world_size, rank):
    """One contiguous band per rank with (nearly) equal total weight; `weights` [H] = expected cost per reference
    row (identical on every rank), e.g. the per-row group counts or device time of an earlier call on similar
    content.  Bands stay non-empty and ordered."""
    w = torch.as_tensor(weights).double().flatten()
    h = int(w.shape[0])
    valid = h - ps + 1
    w = w[:valid] + 1e-3 * float(w[:valid].mean() + 1e-12)     # every row keeps a little weight
    cum = torch.cumsum(w, 0)
    total = float(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(torch.searchsorted(cum, torch.tensor(total * r / world_size, dtype=cum.dtype)).item()) + 1)
    cuts.append(h)
    for r in range(1, world_size + 1):
        cuts[r] = max(cuts[r], cuts[r - 1] + 1) if r < world_size else h
    return cuts[rank], cuts[rank + 1]


def halo_rows(params, max_flow=0.0):
    """Rows above / below a band of reference pixels that its groups can touch: the search window radius plus the
    patch extent plus the drift of the flow trajectory over the temporal search range (both steps)."""
    halo = 0
    for step in (0, 1):
        nwt = max(int(params["sizeSearchTimeFwd"][step]), int(params["sizeSearchTimeBwd"][step]))
        drift = int(math.ceil(nwt * (float(max_flow) + 0.5))) if max_flow > 0 else 0
        halo = max(halo, int(params["sizeSearchWindow"][step]) // 2 + int(params["sizePatch"][step]) - 1 + drift)
    return halo


def make_layout(h, ps, world_size, halo, row_weights=None, margin=0):
    """bands[r] = (y0, y1) reference rows owned by rank r, tiles[r] = (ya, yb) rows rank r holds; global rows.
    `margin` extra rows on each side of a tile let the band borders move by up to `margin` rows during a step
    (rebalance_bands) without any image data moving."""
    if row_weights is None:
        bands = [partition_rows(h, ps, world_size, r) for r in range(world_size)]
    else:
        bands = [partition_rows_weighted(row_weights, ps, world_size, r) for r in range(world_size)]
    tiles = [(max(0, a - halo - margin), min(h, b + halo + margin)) for a, b in bands]
    return bands, tiles


def rebalanced_cuts(remaining_rows, speeds, bands, bands0, margin, ps, damping=1.0):
    """New band borders from a progress report (host arithmetic, identical on every rank).
    remaining_rows [H]: reference pixels still masked per global row; speeds[r]: reference pixels rank r cleared per
    millisecond so far; bands: the current bands (who owns a row now); bands0: the bands the tiles were laid out for.
    A row's remaining time is its count over its owner's speed; the new borders cut the cumulated time into equal
    parts, each border kept within `margin` rows of bands0's border (the image rows a tile holds) and the bands kept
    at least `ps` rows high.  `damping` < 1 moves a border only that fraction of the predicted way (the early speed
    estimate overshoots: measured 2.6 x at N = 2 after two rounds).  Returns the list of new bands."""
    rem = np.asarray(remaining_rows, np.float64)
    h = rem.shape[0]
    world = len(bands)
    tau = np.zeros(h)
    for r, (a, b) in enumerate(bands):
        tau[a:b] = rem[a:b] / max(float(speeds[r]), 1e-9)
    cum = np.cumsum(tau)
    total = float(cum[-1])
    cuts = [0]
    for k in range(1, world):
        orig = bands0[k][0]
        b = int(np.searchsorted(cum, total * k / world)) + 1 if total > 0 else orig
        cur = bands[k][0]
        b = cur + int(round(damping * (b - cur)))          # damping < 1: move only part of the predicted way
        b = min(max(b, orig - margin), orig + margin)
        cuts.append(max(b, cuts[-1] + ps))
    cuts.append(h)
    for k in range(world - 1, 0, -1):                      # keep the last bands non-empty too
        cuts[k] = min(cuts[k], cuts[k + 1] - ps)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _overlap(a, b):
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else None


def exchange_plan(bands, tiles, rank):
    """Row ranges (global) this rank exchanges with every other rank:
    give[s] = rows of MY tile that rank s owns (my accumulations there go to s; s's normalised rows come back),
    take[s] = rows of s's tile that I own      (s's accumulations come to me; my normalised rows go to s)."""
    give, take = {}, {}
    for s in range(len(bands)):
        if s == rank:
            continue
        g = _overlap(tiles[rank], bands[s])
        t = _overlap(tiles[s], bands[rank])
        if g:
            give[s] = g
        if t:
            take[s] = t
    return give, take


# ------------------------------------------------------------------------------------------------ exchanges
def _p2p(sends, recvs, group):
    """sends / recvs: lists of (tensor, peer).  One batched isend/irecv (NCCL groups it; gloo runs it pairwise)."""
    ops = [dist.P2POp(dist.irecv, t, dist.get_global_rank(group, p) if group is not None else p, group) for t, p in recvs]
    ops += [dist.P2POp(dist.isend, t, dist.get_global_rank(group, p) if group is not None else p, group) for t, p in sends]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_accumulators(deno, weights, ya, give, take, group=None):
    """Step 1 of the border exchange: the rows of the tile accumulators (deno [T,C,hb,W], weights [T,hb,W]; tile row 0 =
    global row ya) that another rank owns are sent there; the contributions of the others to my rows are added."""
    t, c, hb, w = deno.shape
    sends, recvs, bufs = [], [], {}
    for s, (lo, hi) in sorted(give.items()):
        buf = torch.empty((t, c + 1, hi - lo, w), dtype=deno.dtype, device=deno.device)
        buf[:, :c] = deno[:, :, lo - ya:hi - ya]
        buf[:, c] = weights[:, lo - ya:hi - ya]
        sends.append((buf, s))
    for s, (lo, hi) in sorted(take.items()):
        bufs[s] = torch.empty((t, c + 1, hi - lo, w), dtype=deno.dtype, device=deno.device)
        recvs.append((bufs[s], s))
    _p2p(sends, recvs, group)
    for s, (lo, hi) in sorted(take.items()):
        deno[:, :, lo - ya:hi - ya] += bufs[s][:, :c]
        weights[:, lo - ya:hi - ya] += bufs[s][:, c]
    return sum(b.numel() for b, _ in sends) * 4


def exchange_halo(img, ya, give, take, group=None):
    """Step 2 of the border exchange: my normalised rows that lie in another rank's tile go there, the normalised rows
    of the others that lie in my tile's halo overwrite my (partial) ones."""
    sends, recvs, bufs = [], [], {}
    for s, (lo, hi) in sorted(take.items()):
        sends.append((img[:, :, lo - ya:hi - ya].contiguous(), s))
    for s, (lo, hi) in sorted(give.items()):
        bufs[s] = torch.empty((img.shape[0], img.shape[1], hi - lo, img.shape[3]), dtype=img.dtype, device=img.device)
        recvs.append((bufs[s], s))
    _p2p(sends, recvs, group)
    for s, (lo, hi) in sorted(give.items()):
        img[:, :, lo - ya:hi - ya] = bufs[s]
    return sum(b.numel() for b, _ in sends) * 4


def gather_bands(img_tile, ya, bands, rank, h, group=None):
    """Every rank contributes the rows of its band; returns the full [T,C,H,W] image on every rank."""
    world = len(bands)
    t, c, _, w = img_tile.shape
    hmax = max(b - a for a, b in bands)
    mine = torch.zeros((t, c, hmax, w), dtype=img_tile.dtype, device=img_tile.device)
    y0, y1 = bands[rank]
    mine[:, :, :y1 - y0] = img_tile[:, :, y0 - ya:y1 - ya]
    allb = torch.empty((world * t, c, hmax, w), dtype=img_tile.dtype, device=img_tile.device)
    dist.all_gather_into_tensor(allb, mine, group=group)
    allb = allb.view(world, t, c, hmax, w)
    full = torch.empty((t, c, h, w), dtype=img_tile.dtype, device=img_tile.device)
    for r, (a, b) in enumerate(bands):
        full[:, :, a:b] = allb[r, :, :, :b - a]
    return full


class BandRebalancer:
    """Mid-step load balancing (the hook of schedule._rounds_async): after the first rounds of a step every rank
    reports how many reference pixels it has cleared per millisecond and how many are left in each row of its band
    (one small all-reduce); the band borders move (rebalanced_cuts, at most `margin` rows: the tiles hold those rows
    already) so that all ranks are predicted to finish together, and the neighbours hand over the mask rows of the
    moved strips in their CURRENT state.  The number of groups a band produces depends on its content (how widely
    the similar patches of a region spread), which is unknown before the step has run for a while."""

    def __init__(self, bands, tiles, rank, group, margin, h, w, t, ps, pt, proc_step, damping=1.0):
        self.damping = float(damping)
        self.bands0 = list(bands)                # borders may move within +-margin of THESE
        self.bands = list(bands)
        self.tiles, self.rank, self.group, self.margin = tiles, rank, group, margin
        self.h, self.ps = h, ps
        self.lattice_per_row = (t - pt + 1) * (w - ps + 1) / float(proc_step ** 2)
        self.t0 = None
        self.moved = []

    def start(self):
        self.t0 = torch.cuda.Event(enable_timing=True)
        self.t0.record()

    def __call__(self, mask):
        world, rank, h = len(self.bands), self.rank, self.h
        (y0, y1), ya = self.bands[rank], self.tiles[rank][0]
        rem = mask.sum(dim=(0, 2), dtype=torch.float32)                       # remaining reference pixels per tile row
        now = torch.cuda.Event(enable_timing=True)
        now.record()
        now.synchronize()
        elapsed = max(self.t0.elapsed_time(now), 1e-3)
        vec = torch.zeros((h + 2 * world,), dtype=torch.float32, device=mask.device)
        vec[y0:y1] = rem[y0 - ya:y1 - ya]
        initial = self.lattice_per_row * max(min(y1, h - self.ps + 1) - y0, 0)
        vec[h + rank] = elapsed
        vec[h + world + rank] = initial
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=self.group)
        v = vec.cpu().numpy().astype(np.float64)
        rem_rows, el, ini = v[:h], v[h:h + world], v[h + world:]
        speeds = [max(ini[r] - rem_rows[a:b].sum(), 1.0) / max(el[r], 1e-3) for r, (a, b) in enumerate(self.bands)]
        new = rebalanced_cuts(rem_rows, speeds, self.bands, self.bands0, self.margin, self.ps, self.damping)
        self._migrate(mask, new)
        self.moved.append(dict(old=self.bands[rank], new=new[rank], elapsed_ms=[round(float(x), 2) for x in el],
                               speeds=[round(float(x), 1) for x in speeds]))
        self.bands = new

    def _migrate(self, mask, new):
        """Mask rows of the strips that change owner go from the old owner to the new one (current state)."""
        rank, ya = self.rank, self.tiles[self.rank][0]
        sends, recvs, put = [], [], []
        for r in range(len(new)):
            if r == rank:
                continue
            lost = _overlap(self.bands[rank], new[r])            # rows I owned that r owns now
            gain = _overlap(self.bands[r], new[rank])            # rows r owned that I own now
            if lost:
                lo, hi = lost
                sends.append((mask[:, lo - ya:hi - ya].contiguous(), r))
            if gain:
                lo, hi = gain
                buf = torch.empty((mask.shape[0], hi - lo, mask.shape[2]), dtype=mask.dtype, device=mask.device)
                recvs.append((buf, r))
                put.append((lo, hi, buf))
        _p2p(sends, recvs, self.group)
        for r in range(len(new)):
            lost = _overlap(self.bands[rank], new[r]) if r != rank else None
            if lost:
                mask[:, lost[0] - ya:lost[1] - ya] = 0
        for lo, hi, buf in put:
            mask[:, lo - ya:hi - ya] = buf


_SIDE_STREAMS = {}


def _side_stream(device):
    """One side stream per device for the whole process.  torch's caching allocator keeps a pool per stream: a stream
    created per call made every call's gather buffers (1.7 GB at 1080p x 30) come from fresh cudaMallocs -- reserved
    memory grew call after call and the cudaMalloc (peer mappings under NCCL) stalled the host for 100-350 ms at
    random (tools/dist_probe.py)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


# ------------------------------------------------------------------------------------------------ inputs
def _rows_to_device(x, ya, yb, device):
    """Rows [ya, yb) of a [T,C,H,W] host (numpy / torch, ideally pinned) or device array as a contiguous float32 device
    tensor.  Host input: one async copy per (t, c) plane -- only these rows cross PCIe."""
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.asarray(x))
    if x.is_cuda:
        return x[:, :, ya:yb].to(device=device, dtype=torch.float32).contiguous()
    t, c, h, w = x.shape
    if x.dtype != torch.float32:
        return x[:, :, ya:yb].to(torch.float32).contiguous().to(device)
    out = torch.empty((t, c, yb - ya, w), dtype=torch.float32, device=device)
    if x.is_contiguous():
        for ti in range(t):
            for ci in range(c):
                out[ti, ci].copy_(x[ti, ci, ya:yb], non_blocking=True)
    else:
        out.copy_(x[:, :, ya:yb])
    return out


def _max_flow_y(flows, y0, y1):
    """max |dy| over the rows of this rank's band (host or device tensors)."""
    m = 0.0
    for key in ("fflow", "bflow"):
        f = flows[key]
        f = f if torch.is_tensor(f) else torch.from_numpy(np.asarray(f))
        m = max(m, float(f[:, 1, y0:y1].abs().max()))
    return m


def denoise_distributed(noisy, sigma, flows=None, schedule="fast", version="default", params=None, stats=None,
                        group=None, device=None, clean=None, max_flow=None, row_weights=None, gather=True,
                        rebalance=None, rebalance_round=None, margin_frac=0.10, rebalance_damping=None):
    """vnlb.denoise over all ranks of `group`.  Every rank passes the same `noisy` [T,C,H,W] (host, ideally pinned, or
    device) and the same `flows`; only its own band + halo rows are copied to its GPU.  Returns (deno, basic, seconds):
    the full frames on every rank (gather=True) or, with gather=False, this rank's band rows only as
    (deno_band, basic_band, ((y0, y1) of deno_band, (y0, y1) of basic_band)) -- the estimate stays sharded like the work.

    max_flow : bound on |flow_y| in pixels per frame used to size the halo; None = measured on this rank's band rows
               (a host pass over them) and agreed over the ranks (max).  The bound is verified on the device.
    row_weights : optional [H] expected cost per reference row for the initial band partition (None = equal row counts).
    rebalance : move the band borders once per step, after `rebalance_round` rounds, to equalise the predicted
               remaining time of the ranks (BandRebalancer); the tiles carry `margin_frac` x band rows of extra
               margin on each side for that.  OFF by default (None = environment VNLB_REBALANCE, else off): measured on
               2 B200s (1920x1080x30, sigma 10) the early speed estimate overshoots (bands 861 k / 808 k groups
               without, 792 k / 877 k after two rounds undamped: 1008 vs 882 ms); with `rebalance_round` = 6 and
               `rebalance_damping` = 0.4 the bands end at 823 k / 845 k but the call is not faster (856 vs 851 ms):
               kept as an option for content whose cost per row is known to be skewed.
    """
    clock = Timer()
    clock.tic()
    import os
    if rebalance is None:                      # experiments: VNLB_REBALANCE=1 [VNLB_REBALANCE_ROUND, VNLB_REBALANCE_DAMP]
        rebalance = os.environ.get("VNLB_REBALANCE", "0") == "1"
    if rebalance_round is None:
        rebalance_round = int(os.environ.get("VNLB_REBALANCE_ROUND", "2"))
    if rebalance_damping is None:
        rebalance_damping = float(os.environ.get("VNLB_REBALANCE_DAMP", "1.0"))
    if schedule != "fast":
        raise ValueError("denoise_distributed runs the throughput schedule (the parity schedule replays the reference's "
                         "single-process random draws and has no multi-GPU meaning)")
    if clean is not None:
        raise ValueError("denoise_distributed: `clean` is not used by the default parameters (srch_img) and is not sharded")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    from .schedule import proc_nl_fast
    T, C, H, W = (int(v) for v in noisy.shape)
    params = params if params is not None else get_params(sigma, False, version)
    ps = int(params["sizePatch"][0])
    with torch.cuda.device(device):
        # ---- layout: band of reference rows, tile = band + halo (+ margin for the moving borders)
        has_flow = flows is not None and flows.get("fflow") is not None and flows.get("bflow") is not None
        if has_flow and max_flow is None:
            y0u, y1u = partition_rows(H, ps, world, rank)
            mf = torch.tensor([_max_flow_y(flows, y0u, y1u)], device=device)
            dist.all_reduce(mf, op=dist.ReduceOp.MAX, group=group)
            max_flow = float(mf.item())
        halo = halo_rows(params, max_flow if has_flow else 0.0)
        rebalance = bool(rebalance) and world > 1
        margin = int(math.ceil(margin_frac * (H / world))) if rebalance else 0
        bands, tiles = make_layout(H, ps, world, halo, row_weights, margin)
        ya, yb = tiles[rank]
        # ---- inputs: only the tile's rows are copied to this GPU
        noisy_t = _rows_to_device(noisy, ya, yb, device)
        dflows = AttrDict(fflow=None, bflow=None)
        flow_ok = None
        if has_flow:
            ff = _rows_to_device(flows["fflow"], ya, yb, device)
            bf = _rows_to_device(flows["bflow"], ya, yb, device)
            ff, bf = expand_flows(dict(fflow=ff, bflow=bf), T)
            dflows.fflow, dflows.bflow = ff.contiguous(), bf.contiguous()
            flow_ok = torch.maximum(ff[:, 1].abs().amax(), bf[:, 1].abs().amax()) <= max_flow + 1e-6
        sent = [0, 0]
        st = stats if stats is not None else {}
        reb = None
        if rebalance:
            reb = BandRebalancer(bands, tiles, rank, group, margin, H, W, T, ps, int(params["sizePatchTime"][0]),
                                 int(params["procStep"][0]), rebalance_damping)
        cur = dict(bands=bands)

        def reduce_fn(images):
            if reb is not None:
                cur["bands"] = reb.bands                   # the borders as they are at the end of the step
            give, take = exchange_plan(cur["bands"], tiles, rank)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if stats is not None else None
            if ev:
                ev[0].record()
            sent[0] += exchange_accumulators(images.deno, images.weights, ya, give, take, group)
            if ev:
                ev[1].record()
                stats.setdefault("exchange_events", []).append(ev)

        def post_fn(images):
            give, take = exchange_plan(cur["bands"], tiles, rank)
            sent[1] += exchange_halo(images.deno, ya, give, take, group)

        noisy_yuv = color.rgb2yuv(noisy_t)
        tile = (ya, H)
        side = _side_stream(device) if gather else None
        outs = []
        basic_yuv = None
        for step in (0, 1):
            y0, y1 = cur["bands"][rank]
            images = alloc.allocate_images_lean(noisy_yuv, basic_yuv, None)
            hook = None
            if reb is not None:
                reb.start()
                hook = (int(rebalance_round), reb)
            proc_nl_fast(images, dflows, get_args(params, C, step, device), st, (y0 - ya, y1 - ya), reduce_fn, tile,
                         post_fn if step == 0 else None, hook)
            band_s = list(cur["bands"])
            if step == 0:
                basic_yuv = images.deno_yuv
            if gather and step == 0:                       # the basic bands travel while step 2 computes
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    outs.append(gather_bands(images.deno, ya, band_s, rank, H, group))
            elif gather:
                outs.append(gather_bands(images.deno, ya, band_s, rank, H, group))
            else:
                b0, b1 = band_s[rank]
                outs.append((images.deno[:, :, b0 - ya:b1 - ya].contiguous(), (b0, b1)))
        if gather:
            torch.cuda.current_stream().wait_stream(side)
            outs[0].record_stream(torch.cuda.current_stream())     # allocated on the side stream, used by the caller on this one
        torch.cuda.synchronize(device)
        if flow_ok is not None and not bool(flow_ok):
            raise ValueError("denoise_distributed: |flow_y| exceeds max_flow = %g on rank %d; the halo (%d rows) is too small"
                             % (max_flow, rank, halo))
        if stats is not None:
            stats["layout"] = dict(band=bands[rank], tile=(ya, yb), halo=halo, margin=margin, rows_copied=yb - ya, rows_total=H)
            stats["exchange_bytes"] = dict(accumulators=sent[0], halo=sent[1])
            if reb is not None:
                stats["rebalance"] = reb.moved
            if "exchange_events" in stats:
                stats["exchange_ms"] = [a.elapsed_time(b) for a, b in stats.pop("exchange_events")]
    if gather:
        return outs[1], outs[0], clock.toc()
    return outs[1][0], outs[0][0], (outs[1][1], outs[0][1])
