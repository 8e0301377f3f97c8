"""One VNLB step: mask -> loop{search, flat, Bayes, aggregate} -> normalise ->
fill holes -> YUV->RGB.  Mirrors lib/vnlb/proc_nl.py:38-141.

Two schedules produce the groups:
  * "parity": the reference's exact schedule -- `nstreams` sequential sub-batches
    of `bsize` pixels drawn with th.randperm on the CPU generator, mask re-read
    between sub-batches (lib/vnlb/search/search.py:38-64).  Same seed => same
    processed pixels as the reference.  Slow by construction (host syncs).
  * "fast": large device-side rounds (see schedule.py); same algorithm and
    parameters, different (still greedy) choice of which masked pixels become
    groups; reported against "parity" by PSNR."""
import torch

from . import agg, alloc, color, deno, search
from . import mask as search_mask
from .flat_areas import update_flat_patch
from .utils import divUp


def batch_params(mask, bsize, nstreams):
    """lib/vnlb/utils/batching.py:10-16."""
    nelems = int(torch.sum(mask).item())
    return nelems, divUp(nelems, nstreams * bsize)


def proc_nl(images, flows, args, stats=None, y_range=None, reduce_fn=None):
    """lib/vnlb/proc_nl.py:38-141 ("parity" schedule).  Mutates images.deno in place."""
    mask, ngroups = search_mask.init_mask(images.shape, args, images.device, y_range)
    patches = alloc.allocate_patches(args.patch_shape, images.clean, images.device)
    bufs = alloc.allocate_bufs(args.bufs_shape, images.device)
    nelems, nbatches = batch_params(mask, args.bsize, args.nstreams)
    color.rgb2yuv_images(images)
    nproc = 0
    for batch in range(nbatches):
        done = search.exec_search(patches, images, flows, mask, bufs, args)
        update_flat_patch(patches, args, bufs.inds)
        nvalid = int(torch.all(bufs.inds != -1, 1).sum().item())
        if nvalid == 0:
            break
        nproc += nvalid
        # get_valid_patches / fill_valid_patches (proc_nl.py:160-177) are folded into
        # the kernels: rows with a -1 index are skipped on the device
        deno.denoise(patches, args, args.deno, bufs.inds)
        agg.agg_patches(patches, images, bufs, args)
        if done:
            break
    finish_step(images, args, reduce_fn)
    if stats is not None:
        stats.setdefault("ngroups", []).append(nproc)
        stats.setdefault("nmask", []).append(nelems)


def finish_step(images, args, reduce_fn=None, post_fn=None):
    """proc_nl.py:118-141: normalise, fill holes, back to RGB.  Multi-GPU hooks: `reduce_fn(images)` adds the other
    ranks' contributions to the accumulators first; `post_fn(images)` runs on the normalised YUV estimate (exchange of
    the halo rows the next step searches).  Lean image sets (images.is_yuv: the throughput schedule's, whose inputs
    were converted once by the caller) convert only the estimate back; images.deno_yuv keeps its YUV version."""
    if reduce_fn is not None:
        reduce_fn(images)
    agg.normalize(images, args)
    if post_fn is not None:
        post_fn(images)
    if images.get("is_yuv"):
        images.deno_yuv = images.deno
        images.deno = color.yuv2rgb_new(images.deno_yuv)
    else:
        color.yuv2rgb_images(images)
        torch.cuda.synchronize(images.device)
