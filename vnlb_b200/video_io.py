"""Frame-sequence IO around `denoise`: mirrors
lib/vnlb/utils/video_io.py:14-37 (read_video_sequence) and the PSNR / relative-error
report of scripts/compare_cpp.py:25-58.  Host-side convenience, not on the hot path."""
import glob
import os

import numpy as np


def read_video_sequence(folder, max_frames=None, ext=None):
    """Frames `folder/*.{png,jpg,tif,npy}` in name order -> float32 [T,3,H,W], RGB in 0..255."""
    import cv2
    exts = [ext] if ext else ["png", "jpg", "jpeg", "tif", "tiff", "npy"]
    files = sorted(f for e in exts for f in glob.glob(os.path.join(folder, "*." + e)))
    if max_frames:
        files = files[:max_frames]
    if not files:
        raise FileNotFoundError("no frames found in %s" % folder)
    frames = []
    for f in files:
        if f.endswith(".npy"):
            fr = np.load(f).astype(np.float32)
            if fr.ndim == 3 and fr.shape[-1] in (1, 3):
                fr = fr.transpose(2, 0, 1)
        else:
            img = cv2.imread(f, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise IOError("cannot read %s" % f)
            if img.ndim == 2:
                img = img[..., None].repeat(3, -1)
            else:
                img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
            fr = np.transpose(img, (2, 0, 1)).astype(np.float32)
        frames.append(fr)
    return np.ascontiguousarray(np.stack(frames))


def save_video_sequence(vid, folder, fmt="%05d.png"):
    """float [T,3,H,W] in 0..255 -> 8-bit frames (plus .npy when fmt ends with npy)."""
    import cv2
    os.makedirs(folder, exist_ok=True)
    vid = np.asarray(vid)
    for t in range(vid.shape[0]):
        path = os.path.join(folder, fmt % t)
        if path.endswith(".npy"):
            np.save(path, vid[t])
        else:
            img = np.clip(np.rint(vid[t].transpose(1, 2, 0)), 0, 255).astype(np.uint8)
            cv2.imwrite(path, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))


def compare_report(ours, other, clean=None):
    """scripts/compare_cpp.py:37-58: mean relative error and PSNR difference between two outputs."""
    from .utils import compute_psnrs
    ours, other = np.asarray(ours, np.float64), np.asarray(other, np.float64)
    rep = {"Ave Rel. Error": float(np.mean(np.abs(other - ours) / (np.abs(other) + 1e-10)))}
    if clean is not None:
        a = float(compute_psnrs(other, clean).mean())
        b = float(compute_psnrs(ours, clean).mean())
        rep.update({"other_psnr": a, "our_psnr": b, "Abs. Error (PSNR)": abs(a - b), "Rel. Error (PSNR)": abs(a - b) / abs(a)})
    return rep
