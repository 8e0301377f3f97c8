"""Small host helpers (edict-like container, flow handling, timing, PSNR)."""
import time

import numpy as np
import torch


class AttrDict(dict):
    """Attribute-access dict standing in for the reference's EasyDict usage."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def divUp(a, b):
    return (a - 1) // b + 1


class Timer:
    """lib/vnlb/utils/timer.py:9-42 (wall clock; toc() returns the running average)."""

    def __init__(self):
        self.total_time, self.calls, self.start_time, self.diff, self.average_time = 0., 0, 0., 0., 0.

    def tic(self):
        self.start_time = time.time()

    def toc(self, average=True):
        self.diff = time.time() - self.start_time
        self.total_time += self.diff
        self.calls += 1
        self.average_time = self.total_time / self.calls
        return self.average_time if average else self.diff


def compute_psnrs(deno, clean, imax=255.):
    """lib/vnlb/utils/metrics.py:50-71: per-frame PSNR, peak `imax`."""
    if torch.is_tensor(deno):
        deno = deno.detach().cpu().numpy()
    if torch.is_tensor(clean):
        clean = clean.detach().cpu().numpy()
    d = np.asarray(deno, np.float64) / imax
    c = np.asarray(clean, np.float64) / imax
    return -10 * np.log10(((d - c) ** 2).mean(axis=(-3, -2, -1)))


def expand_flows(flows, t):
    """check_and_expand_flows / expand_flows_th, lib/vnlb/utils/utils.py:24-34,143-158:
    [T-1,2,H,W] flows are expanded to T frames by repeating the last forward and
    the first backward flow (the C++ convention)."""
    fflow, bflow = flows["fflow"], flows["bflow"]
    if fflow.shape[0] != bflow.shape[0]:
        raise ValueError("num flows must be equal.")
    if fflow.shape[0] == t - 1:
        fflow = torch.cat([fflow, fflow[[-1]]], dim=0)
        bflow = torch.cat([bflow[[0]], bflow], dim=0)
    elif fflow.shape[0] != t:
        raise ValueError("The input flows are the wrong shape.\n(nframes,two,height,width)")
    return fflow, bflow


def prepare_flows(flows, shape, device):
    """None / dict(fflow, bflow) of numpy or torch -> AttrDict of contiguous f32
    CUDA tensors [T,2,H,W], or fflow = bflow = None for zero flow (what the
    reference always uses: lib/vnlb/alloc.py:66-72)."""
    out = AttrDict(fflow=None, bflow=None)
    if flows is None:
        return out
    t, c, h, w = shape
    ff, bf = flows["fflow"], flows["bflow"]
    if ff is None or bf is None:
        return out
    ff = torch.as_tensor(ff).to(device=device, dtype=torch.float32)
    bf = torch.as_tensor(bf).to(device=device, dtype=torch.float32)
    ff, bf = expand_flows(dict(fflow=ff, bflow=bf), t)
    if tuple(ff.shape) != (t, 2, h, w) or tuple(bf.shape) != (t, 2, h, w):
        raise ValueError("flows must have shape (nframes,2,height,width)")
    out.fflow, out.bflow = ff.contiguous(), bf.contiguous()
    return out


def read_flo(path):
    """Middlebury .flo reader (lib/vnlb/utils/flow_utils.py:14-63): returns [2,H,W]
    float32 with channel 0 = u (x), channel 1 = v (y)."""
    with open(path, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != 202021.25:
            raise ValueError("Magic number incorrect. Invalid .flo file")
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
    return np.ascontiguousarray(data.reshape(h, w, 2).transpose(2, 0, 1))
