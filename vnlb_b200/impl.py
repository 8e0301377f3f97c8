"""Public API: the two-step VNLB driver.  Mirrors lib/vnlb/impl.py:24-62."""
import numpy as np
import torch

from . import alloc
from .params import get_args, get_params
from .proc_nl import proc_nl
from .utils import Timer, prepare_flows


def denoise(noisy, sigma, gpuid=0, clean=None, verbose=True, flows=None, schedule="fast",
            version="default", params=None, stats=None):
    """Video Non-Local Bayes (VNLB).

    Drop-in for `vnlb.denoise(noisy, sigma, gpuid=0, clean=None, verbose=True)`
    (lib/vnlb/impl.py:24) plus the north-star `flows=` argument.

    noisy : [T,C,H,W] RGB in 0..255, numpy or torch (any float dtype; computed in float32)
    sigma : noise standard deviation in 0..255 units
    flows : None (zero flow, what the reference always uses) or a dict with
            'fflow'/'bflow' of shape [T,2,H,W] or [T-1,2,H,W] (channel 0 = dx, 1 = dy)
    schedule : "fast" (device-side rounds) or "parity" (the reference's exact
            sub-batch schedule and th.randperm draws)
    returns (deno, basic, seconds): float32 CUDA tensors [T,C,H,W] and the wall time
            of the call including the host->device copy, after a device sync.
    """
    clock = Timer()
    clock.tic()
    if not torch.cuda.is_available():
        raise RuntimeError("vnlb_b200.denoise needs a CUDA device (B200, sm_100a); there is no CPU path")
    if gpuid < 0:
        raise ValueError("gpuid must name a CUDA device (the reference's gpuid=-1 CPU mode does not exist here)")
    device = torch.device("cuda:%d" % gpuid)
    if schedule not in ("fast", "parity"):
        raise ValueError("unknown schedule [%s]" % schedule)
    with torch.cuda.device(device):
        if not torch.is_tensor(noisy):
            noisy = torch.from_numpy(np.ascontiguousarray(noisy))
        noisy = noisy.to(device=device, dtype=torch.float32).contiguous()
        if noisy.dim() != 4:
            raise ValueError("noisy must be [T,C,H,W]")
        if clean is not None and not torch.is_tensor(clean):
            clean = torch.from_numpy(np.ascontiguousarray(clean))
        if clean is not None:
            clean = clean.to(device=device, dtype=torch.float32).contiguous()
        c = noisy.shape[1]
        params = params if params is not None else get_params(sigma, verbose, version)
        dflows = prepare_flows(flows, noisy.shape, device)
        step_fn = proc_nl
        if schedule == "fast":
            from .schedule import proc_nl_fast
            step_fn = proc_nl_fast

        # -- [step 1] --
        images = alloc.allocate_images(noisy, None, clean)
        args = get_args(params, c, 0, device)
        step_fn(images, dflows, args, stats)
        basic = images["deno"].clone()

        # -- [step 2] --
        images = alloc.allocate_images(noisy, basic, clean)
        args = get_args(params, c, 1, device)
        step_fn(images, dflows, args, stats)
        deno = images["deno"]
        torch.cuda.synchronize(device)
    return deno, basic, clock.toc()
