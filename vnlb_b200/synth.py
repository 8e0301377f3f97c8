"""Seeded synthetic video for benchmarks and smoke runs (no datasets offline):
smooth low-frequency field + textured rectangles translating 0-2 px/frame,
range 0..255, float32 [T,C,H,W]; noisy = clean + N(0, sigma^2), unclipped.
return_flows=True also returns the analytic forward / backward flows [T,2,H,W] (channel 0 = dx, 1 = dy; the objects'
translations, zero on the background)."""
import numpy as np


def synth_video(T, H, W, seed=123, C=3, return_flows=False, crop=None):
    rng = np.random.RandomState(seed)
    # crop = (Hc, Wc): only the top-left Hc x Wc corner of the H x W video is materialised (same content as cropping
    # the full video; the CPU baseline's bounded sample of a 1080p workload)
    Hc, Wc = (H, W) if crop is None else (min(int(crop[0]), H), min(int(crop[1]), W))
    yy, xx = np.meshgrid(np.arange(Hc, dtype=np.float32), np.arange(Wc, dtype=np.float32), indexing="ij")
    vid = np.zeros((T, C, Hc, Wc), np.float32)
    ff = np.zeros((T, 2, Hc, Wc), np.float32) if return_flows else None
    bf = np.zeros((T, 2, Hc, Wc), np.float32) if return_flows else None
    rects = []
    for _ in range(6):
        rh, rw = rng.randint(H // 6 + 2, H // 2 + 3), rng.randint(W // 6 + 2, W // 2 + 3)
        y0, x0 = rng.randint(0, max(1, H - rh)), rng.randint(0, max(1, W - rw))
        vy, vx = rng.randint(-2, 3), rng.randint(-2, 3)
        tex = rng.rand(C, rh, rw).astype(np.float32) * 60 + rng.rand(C, 1, 1).astype(np.float32) * 150
        fy, fx = rng.uniform(0.2, 1.2), rng.uniform(0.2, 1.2)
        stripes = 25 * np.sin(fy * np.arange(rh)[:, None] + fx * np.arange(rw)[None, :]).astype(np.float32)
        rects.append((y0, x0, rh, rw, vy, vx, tex * 0.3 + stripes[None] + 60))
    for t in range(T):
        for ch in range(C):
            vid[t, ch] = 110 + 60 * np.sin(xx / (17. + 3 * ch) + 0.1 * t) * np.cos(yy / (23. - 2 * ch))
        for (y0, x0, rh, rw, vy, vx, tex) in rects:
            ya, xa = y0 + vy * t, x0 + vx * t
            ys, xs = max(0, ya), max(0, xa)
            ye, xe = min(Hc, ya + rh), min(Wc, xa + rw)
            if ye > ys and xe > xs:
                vid[t, :, ys:ye, xs:xe] = tex[:, ys - ya:ye - ya, xs - xa:xe - xa]
                if return_flows:     # the object's own translation (the last-drawn object wins, as in the frame)
                    ff[t, 0, ys:ye, xs:xe], ff[t, 1, ys:ye, xs:xe] = vx, vy
                    bf[t, 0, ys:ye, xs:xe], bf[t, 1, ys:ye, xs:xe] = -vx, -vy
    vid = np.clip(vid, 0, 255).astype(np.float32)
    if return_flows:     # "precomputed flows" of SURVEY 8d: the generator's analytic translation field, |flow| <= 2 px/frame
        return vid, dict(fflow=ff, bflow=bf)
    return vid


def add_noise(clean, sigma, seed=123):
    rng = np.random.RandomState(seed + 1)
    return (clean + rng.randn(*clean.shape).astype(np.float32) * np.float32(sigma)).astype(np.float32)
