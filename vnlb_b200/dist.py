"""Multi-GPU VNLB: one process per GPU (torch.distributed, NCCL over NVLink).

The path shards over REFERENCE PIXELS: every rank holds the whole
video (a 1080p x 30 float32 video is 746 MB, trivial next to 180 GB of HBM) and
owns one horizontal band of the reference-pixel lattice; it searches, filters
and aggregates only the groups whose reference pixel lies in its band.  Groups
of neighbouring bands overlap at the band borders (patches reach a search-window
radius into the neighbour), so the per-step accumulators (sum image + weights)
are summed across ranks -- the one collective of the path -- and every rank then
normalises locally, which also gives each rank the complete `basic` image that
step 2 searches.  Each rank runs its own greedy mask over its own band, so the
multi-GPU output matches the single-GPU output within the PSNR tolerance
(0.02 dB), not max-abs."""
import numpy as np
import torch
import torch.distributed as dist

from . import alloc
from .params import get_args, get_params
from .utils import Timer, prepare_flows


def partition_rows(h, ps, world_size, rank):
    """Band [y0, y1) of reference-pixel rows (valid rows are 0..h-ps) of `rank`."""
    valid = h - ps + 1
    y0 = (valid * rank) // world_size
    y1 = (valid * (rank + 1)) // world_size
    if rank == world_size - 1:
        y1 = h
    return y0, y1


def partition_rows_snake(h, ps, world_size, rank):
    """Two bands per rank, assigned boustrophedon (rank r owns bands r and 2N-1-r of 2N), so that
    content that changes from the top to the bottom of the frame is averaged over the ranks:
    the number of groups a band produces depends on its texture, not only on its size."""
    if world_size == 1:
        return [partition_rows(h, ps, 1, 0)]
    a = partition_rows(h, ps, 2 * world_size, rank)
    b = partition_rows(h, ps, 2 * world_size, 2 * world_size - 1 - rank)
    return [a, b]


def partition_rows_weighted(weights, ps, world_size, rank):
    """One contiguous band per rank with (nearly) equal total weight; `weights` [H] = groups per reference
    row observed in the previous step (identical on every rank).  Fewer band borders than the snake
    partition and balanced on the actual content."""
    h = int(weights.shape[0])
    valid = h - ps + 1
    w = weights[:valid].double() + 1e-3                    # every row keeps a little weight
    cum = torch.cumsum(w, 0)
    total = float(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(torch.searchsorted(cum, torch.tensor(total * r / world_size, dtype=cum.dtype,
                                                             device=cum.device)).item()) + 1)
    cuts.append(h)
    for r in range(1, world_size + 1):                     # keep the bands non-empty and ordered
        cuts[r] = max(cuts[r], cuts[r - 1] + 1) if r < world_size else h
    return cuts[rank], cuts[rank + 1]


def allreduce_accumulators(images, group=None):
    """Sum the aggregation accumulators over ranks (border overlap + band union)."""
    dist.all_reduce(images.deno, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(images.weights, op=dist.ReduceOp.SUM, group=group)


def denoise_distributed(noisy, sigma, flows=None, schedule="fast", version="default", params=None,
                        stats=None, group=None, device=None, clean=None, balance=True):
    """vnlb.denoise over all ranks of `group`.  Every rank passes the same `noisy`
    (host or device) and receives the full (deno, basic, seconds).

    balance: True / "snake" = two boustrophedon bands per rank in both steps (default; measured best at
    N = 8: 394.6 ms per call); "weighted" = step-2 bands equalised on the step-1 group histogram
    (measured worse: 405.4 ms -- the cost of a group in step 2 depends on its content, not only on the
    count); False = one plain band per rank."""
    clock = Timer()
    clock.tic()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    from .proc_nl import proc_nl
    from .schedule import proc_nl_fast
    step_fn = proc_nl_fast if schedule == "fast" else proc_nl
    with torch.cuda.device(device):
        if not torch.is_tensor(noisy):
            noisy = torch.from_numpy(np.ascontiguousarray(noisy))
        noisy = noisy.to(device=device, dtype=torch.float32).contiguous()
        t, c, h, w = noisy.shape
        params = params if params is not None else get_params(sigma, False, version)
        dflows = prepare_flows(flows, noisy.shape, device)

        def reduce_fn(images):
            if stats is None:
                allreduce_accumulators(images, group)
                return
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            allreduce_accumulators(images, group)
            e1.record()
            stats.setdefault("allreduce_events", []).append((e0, e1))

        basic = None
        row_hist = None
        st = stats if stats is not None else {}
        for step in (0, 1):
            images = alloc.allocate_images(noisy, basic, clean)
            args = get_args(params, c, step, device)
            mode = balance
            if mode == "auto" or mode is True:
                mode = "snake"
            if not mode:
                y_range = partition_rows(h, args.ps, world, rank)
            elif step == 1 and row_hist is not None and mode == "weighted":
                y_range = partition_rows_weighted(row_hist, args.ps, world, rank)     # balanced on step-1 group density
            else:
                y_range = partition_rows_snake(h, args.ps, world, rank)
            st["want_row_hist"] = bool(mode == "weighted" and step == 0 and schedule == "fast")
            step_fn(images, dflows, args, st, y_range, reduce_fn)
            if step == 0:
                basic = images["deno"].clone()
                if st.get("row_hist") is not None:
                    row_hist = st.pop("row_hist")
                    dist.all_reduce(row_hist, op=dist.ReduceOp.SUM, group=group)
        st.pop("want_row_hist", None)
        deno = images["deno"]
        torch.cuda.synchronize(device)
        if stats is not None and "allreduce_events" in stats:
            stats["allreduce_ms"] = [a.elapsed_time(b) for a, b in stats.pop("allreduce_events")]
    return deno, basic, clock.toc()
