"""Pre-allocated image set, patch stacks and top-k buffers.
Mirrors lib/vnlb/alloc.py (allocate_patches :10-30, allocate_images :32-64,
allocate_flows :66-72, allocate_bufs :74-87).  Everything is float32 on the
device: the reference keeps the caller's dtype for images (float64 in its
README example); this implementation casts to float32 at entry."""
import torch

from .utils import AttrDict


def allocate_patches(shape, clean, device):
    tsize, npa, ps_t, c, ps, ps = shape
    patches = AttrDict()
    patches.noisy = torch.zeros((tsize, npa, ps_t, c, ps, ps), dtype=torch.float32, device=device)
    patches.basic = torch.zeros((tsize, npa, ps_t, c, ps, ps), dtype=torch.float32, device=device)
    patches.clean = None
    if clean is not None:
        patches.clean = torch.zeros((tsize, npa, ps_t, c, ps, ps), dtype=torch.float32, device=device)
    patches.flat = torch.zeros((tsize,), dtype=torch.uint8, device=device)
    patches.shape = list(shape)
    patches.images = ["noisy", "basic", "clean"]
    patches.tensors = ["noisy", "basic", "clean", "flat"]
    return patches


def allocate_images(noisy, basic, clean):
    imgs = AttrDict()
    imgs.noisy = noisy
    imgs.shape = noisy.shape
    imgs.device = noisy.device
    t, c, h, w = noisy.shape
    imgs.basic = basic
    if basic is None:
        imgs.basic = torch.zeros((t, c, h, w), dtype=torch.float32, device=noisy.device)
    imgs.clean = clean
    if clean is not None and not torch.is_tensor(clean):
        imgs.clean = torch.from_numpy(clean).to(device=noisy.device, dtype=torch.float32)
    imgs.deno = torch.zeros((t, c, h, w), dtype=torch.float32, device=noisy.device)
    imgs.weights = torch.zeros((t, h, w), dtype=torch.float32, device=noisy.device)
    imgs.vals = torch.zeros((t, h, w), dtype=torch.float32, device=noisy.device)
    imgs.patch_images = ["noisy", "basic", "clean"]
    imgs.ikeys = ["noisy", "basic", "clean", "deno"]
    return imgs


def allocate_images_lean(noisy_yuv, basic_yuv, clean_yuv=None):
    """Image set of the throughput schedule: the inputs are ALREADY in YUV (converted once per call by the caller, not
    once per step), `basic` stays None in step 1 (the reference gathers an all-zero image there, alloc.py:45-50), and
    only the accumulators are zero-filled -- no `vals` image (never read: comp_agg.py:140-141)."""
    imgs = AttrDict()
    imgs.noisy, imgs.basic, imgs.clean = noisy_yuv, basic_yuv, clean_yuv
    imgs.shape, imgs.device = noisy_yuv.shape, noisy_yuv.device
    t, c, h, w = noisy_yuv.shape
    imgs.deno = torch.zeros((t, c, h, w), dtype=torch.float32, device=noisy_yuv.device)
    imgs.weights = torch.zeros((t, h, w), dtype=torch.float32, device=noisy_yuv.device)
    imgs.vals = None
    imgs.is_yuv = True
    imgs.patch_images = ["noisy", "basic", "clean"]
    imgs.ikeys = ["noisy", "basic", "clean", "deno"]
    return imgs


def allocate_flows(shape, device):
    """Zero flows.  The search treats fflow = bflow = None as zero flow, so no
    [T,2,H,W] tensors are materialised (reference: alloc.py:66-72)."""
    return AttrDict(fflow=None, bflow=None)


def allocate_bufs(shape, device):
    tsize, npa = shape
    bufs = AttrDict()
    bufs.vals = torch.zeros((tsize, npa), dtype=torch.float32, device=device)
    bufs.inds = -torch.ones((tsize, npa), dtype=torch.int64, device=device)
    bufs.shape = list(shape)
    return bufs
