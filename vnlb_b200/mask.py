"""Reference-pixel mask: lattice initialisation, random draw of reference
pixels and the greedy clearing of found neighbours ("paste trick" + aggregation
boost).  Mirrors lib/vnlb/search_mask/mask.py: init_mask :190-213, mask2inds
:18-31, update_mask :33-35, update_mask_inds :37-86, agg_boost :104-187."""
import torch

from . import _lib as L


def init_mask(shape, args, device=None, y_range=None):
    """mask.py:190-213 (+ comp_params :252-288, fill_mask :315-358), computed on
    the device.  Returns (mask int8 [T,H,W], ngroups).  `y_range` = (y0, y1)
    restricts the set pixels to a row band (multi-GPU partition)."""
    t, c, h, w = shape
    device = device if device is not None else args.device
    mask = init_mask_device(shape, args, device, y_range)
    return mask, int(mask.sum().item())


def init_mask_device(shape, args, device, y_range=None, tile=None):
    """The mask only (no host sync).  `y_range`: None, (y0, y1) or a list of such bands (local rows).
    `tile` = (y_offset, H_total): `shape` is a row tile of a frame of H_total rows starting at global row
    y_offset (multi-GPU band + halo); the lattice follows the global rows."""
    t, c, h, w = shape
    bands = [(0, h)] if y_range is None else ([y_range] if isinstance(y_range[0], int) else list(y_range))
    y_off, h_total = (0, h) if tile is None else (int(tile[0]), int(tile[1]))
    mask = None
    for (y0, y1) in bands:
        m = torch.empty((t, h, w), dtype=torch.int8, device=device)
        L.check(L.lib.vnlb_init_mask_tile(L.ptr(m), t, h, w, args.ps, args.pt, args.procStep, int(y0), int(y1),
                                          y_off, h_total, L.stream_ptr()), "vnlb_init_mask_tile")
        mask = m if mask is None else torch.bitwise_or(mask, m)
    return mask


def mask2inds(mask, bsize, rand=True, order=None):
    """mask.py:18-31: up to `bsize` random set pixels as (t,y,x) rows.  The
    permutation is drawn with th.randperm on the CPU default generator exactly as
    the reference does, so a seeded run visits the same pixels."""
    index = torch.nonzero(mask)
    if index.shape[0] == 0:
        return index
    if rand:
        mlen = max(len(index), bsize)
        if order is None or mlen == bsize:
            order = torch.randperm(index.shape[0])
        return index[order[:bsize].to(index.device)]
    return index[:bsize]


def update_mask(mask, access, val=0):
    """mask.py:33-35."""
    assert access.shape[1] == 3
    mask[access[:, 0], access[:, 1], access[:, 2]] = val


def update_mask_inds(mask, inds, chnls, cs_ptr=None, boost=True, val=0, nkeep=-1):
    """mask.py:37-86: clear the mask at every index of every valid row of `inds`
    and, with `boost`, at the 4 spatial neighbours (agg_boost :104-187)."""
    if val != 0:
        raise ValueError("update_mask_inds: only val=0 is supported")
    if nkeep != -1:
        inds = inds[:, :nkeep].contiguous()
    t, h, w = mask.shape
    b, k = inds.shape
    if b == 0:
        return
    L.check(L.lib.vnlb_mask_update(L.ptr(mask, torch.int8), L.ptr(inds, torch.int64), b, k, t, chnls, h, w,
                                   int(bool(boost)), L.stream_ptr(cs_ptr)), "vnlb_mask_update")
