"""Public API: the two-step VNLB driver.  Mirrors lib/vnlb/impl.py:24-62."""
import numpy as np
import torch

from . import alloc, color
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
    version  : parameter table, see params.get_params.  The default is the classic VNLB table (`default_params`,
            what BASELINE.json names); the reference's own get_params hard-codes "iphone" (15x15 window, +-10
            frames, needle search), so `vnlb.denoise(noisy, sigma)` corresponds to version="iphone" here (served
            with the l2 search).  The similarity search follows the restated semantics documented in README.md
            (window_mode, dist_chnls): the reference's search lives in the absent vpss package.
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
        from .schedule import proc_nl_fast

        if schedule == "fast":
            # throughput schedule: colour conversion once per call (not once per step and image), no zero images
            # converted, no clone of the basic estimate
            noisy_yuv = color.rgb2yuv(noisy)
            clean_yuv = color.rgb2yuv(clean) if clean is not None else None
            images = alloc.allocate_images_lean(noisy_yuv, None, clean_yuv)
            proc_nl_fast(images, dflows, get_args(params, c, 0, device), stats)
            basic, basic_yuv = images.deno, images.deno_yuv
            images = alloc.allocate_images_lean(noisy_yuv, basic_yuv, clean_yuv)
            proc_nl_fast(images, dflows, get_args(params, c, 1, device), stats)
            deno = images.deno
        else:
            # -- [step 1] --
            images = alloc.allocate_images(noisy, None, clean)
            args = get_args(params, c, 0, device)
            proc_nl(images, dflows, args, stats)
            basic = images["deno"].clone()

            # -- [step 2] --
            images = alloc.allocate_images(noisy, basic, clean)
            args = get_args(params, c, 1, device)
            proc_nl(images, dflows, args, stats)
            deno = images["deno"]
        torch.cuda.synchronize(device)
    return deno, basic, clock.toc()
