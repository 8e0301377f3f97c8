"""Properties and an independent brute-force check of the CPU search oracle
(the restatement of the absent `vpss`; parity unpinned, see oracle header)."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import vnlb_oracle as orc


def sargs(ps=7, pt=2, w_s=27, f=6, b=6, k=100, step=0, c=3):
    return SimpleNamespace(ps=ps, pt=pt, w_s=w_s, nWt_f=f, nWt_b=b, npatches=k, step=step, c=c)


def rand_img(T, C, H, W, seed=0):
    return (np.random.RandomState(seed).rand(T, C, H, W) * 255).astype(np.float32)


def brute(img, t0, y0, x0, a, dc, ff=None, bf=None):
    """Independent float64 restatement of the window logic + distances."""
    T, C, H, W = img.shape
    shift = min(0, t0 - a.nWt_b) + max(0, t0 + a.nWt_f - T + a.pt)
    r0, r1 = max(0, t0 - a.nWt_b - shift), min(T - a.pt, t0 + a.nWt_f - shift)
    cx, cy = {t0: x0}, {t0: y0}

    def rnd(v):  # roundf: half away from zero
        return int(np.floor(abs(v) + 0.5) * np.sign(v))
    for qt in range(t0 + 1, r1 + 1):
        dx = ff[qt - 1, 0, cy[qt - 1], cx[qt - 1]] if ff is not None else 0.
        dy = ff[qt - 1, 1, cy[qt - 1], cx[qt - 1]] if ff is not None else 0.
        cx[qt] = min(max(rnd(np.float32(cx[qt - 1]) + np.float32(dx)), 0), W - 1)
        cy[qt] = min(max(rnd(np.float32(cy[qt - 1]) + np.float32(dy)), 0), H - 1)
    for qt in range(t0 - 1, r0 - 1, -1):
        dx = bf[qt + 1, 0, cy[qt + 1], cx[qt + 1]] if bf is not None else 0.
        dy = bf[qt + 1, 1, cy[qt + 1], cx[qt + 1]] if bf is not None else 0.
        cx[qt] = min(max(rnd(np.float32(cx[qt + 1]) + np.float32(dx)), 0), W - 1)
        cy[qt] = min(max(rnd(np.float32(cy[qt + 1]) + np.float32(dy)), 0), H - 1)
    half = (a.w_s - 1) // 2
    ref = img[t0:t0 + a.pt, :dc, y0:y0 + a.ps, x0:x0 + a.ps].astype(np.float64)
    out = []
    for qt in range(r0, r1 + 1):
        def rng(c, L):
            s = min(0, c - half) + max(0, c + half - L + a.ps)
            return max(0, c - half - s), min(L - a.ps, c + half - s)
        xa, xb = rng(cx[qt], W)
        ya, yb = rng(cy[qt], H)
        for qy in range(ya, yb + 1):
            for qx in range(xa, xb + 1):
                cand = img[qt:qt + a.pt, :dc, qy:qy + a.ps, qx:qx + a.ps].astype(np.float64)
                out.append((((ref - cand) ** 2).sum(), qt * C * H * W + qy * W + qx))
    return out


@pytest.mark.parametrize("cfg", [
    dict(shape=(3, 3, 32, 32), a=sargs(), step=0),                  # reference test's random case
    dict(shape=(3, 3, 32, 32), a=sargs(ps=3), step=1),              # {'ps_x':3,'ps_t':2}
    dict(shape=(5, 3, 24, 40), a=sargs(w_s=9, f=2, b=2, k=20), step=1),
    dict(shape=(4, 1, 20, 20), a=sargs(ps=5, pt=1, w_s=7, f=1, b=1, k=10, c=1), step=0),
])
def test_search_vs_bruteforce(cfg):
    T, C, H, W = cfg["shape"]
    a = cfg["a"]
    a.step = cfg["step"]
    a.c = C
    img = rand_img(T, C, H, W, 1)
    rs = np.random.RandomState(2)
    q = np.stack([rs.randint(0, T - a.pt + 1, 12), rs.randint(0, H - a.ps + 1, 12),
                  rs.randint(0, W - a.ps + 1, 12)], 1).astype(np.int64)
    q[0] = (0, 0, 0)
    q[1] = (T - a.pt, H - a.ps, W - a.ps)
    k = a.npatches
    vals = np.full((16, k), np.inf, np.float32)
    inds = np.full((16, k), -1, np.int64)
    orc.exec_sim_search_burst(img, q, vals, inds, None, 20., a)
    dc = 1 if a.step == 0 else C
    for i, (t0, y0, x0) in enumerate(q):
        ref = sorted(brute(img, t0, y0, x0, a, dc))
        m = min(k, len(ref))
        assert np.all(inds[i, :m] >= 0) and np.all(inds[i, m:] == -1)
        assert inds[i, 0] == t0 * C * H * W + y0 * W + x0 and vals[i, 0] == 0.      # self match first
        assert np.all(np.diff(vals[i, :m]) >= 0)                                     # ascending
        np.testing.assert_allclose(vals[i, :m], [r[0] for r in ref[:m]], rtol=2e-6)
        assert set(inds[i, :m]) == set(r[1] for r in ref[:m])                        # random data: no ties
    # rows beyond Q keep the sentinels
    assert np.all(inds[12:] == -1) and np.all(np.isinf(vals[12:]))


def test_window_keeps_size_at_borders():
    """shift mode: a corner query on a big-enough image still sees w_s^2 x nfr candidates."""
    a = sargs(w_s=9, f=2, b=2, k=5)
    img = rand_img(8, 3, 40, 40, 3)
    d, ind = orc.search_all(img, 0, 0, 0, None, a)
    assert d.shape[0] == 9 * 9 * 5
    d, ind = orc.search_all(img, 6, 33, 33, None, a)           # last valid corner
    assert d.shape[0] == 9 * 9 * 5
    d2, _ = orc.search_all(img, 6, 33, 33, None, a, window_mode="clip")
    assert d2.shape[0] == 5 * 5 * 3


def test_small_image_invalid_rows():
    """fewer candidates than k -> row keeps -1 and is invalid (proc_nl.py:167)."""
    a = sargs(k=100)
    img = rand_img(2, 3, 9, 9, 4)                                 # 3x3x1 = 9 candidates
    vals = np.full((1, 100), np.inf, np.float32)
    inds = np.full((1, 100), -1, np.int64)
    orc.exec_sim_search_burst(img, np.array([[0, 1, 1]]), vals, inds, None, 20., a)
    assert (inds[0] >= 0).sum() == 9 and inds[0, 9] == -1


def test_flow_trajectory():
    """A pattern translating by (+3,-2) px/frame is found at distance 0 along the flow."""
    T, C, H, W = 6, 3, 48, 48                                     # valid start frames 0..4
    base = rand_img(1, C, H + 40, W + 40, 5)[0]
    vid = np.stack([base[:, 20 - (-2) * t:20 - (-2) * t + H, 20 - 3 * t:20 - 3 * t + W] for t in range(T)])
    ff = np.zeros((T, 2, H, W), np.float32)
    ff[:, 0], ff[:, 1] = 3.4, -1.6                                # rounds to (+3,-2)
    bf = -ff
    a = sargs(w_s=5, f=2, b=2, k=5, step=1)
    vals = np.full((1, 5), np.inf, np.float32)
    inds = np.full((1, 5), -1, np.int64)
    orc.exec_sim_search_burst(vid, np.array([[2, 20, 20]]), vals, inds, dict(fflow=ff, bflow=bf), 20., a)
    assert np.all(vals[0] == 0)
    chw, hw = C * H * W, H * W
    got = sorted((int(i // chw), int((i % hw) // W), int(i % W)) for i in inds[0])
    assert got == [(t, 20 - 2 * (t - 2), 20 + 3 * (t - 2)) for t in range(5)]
    # ties (all zero) are ordered by enumeration order: frame ascending
    assert list(inds[0] // chw) == [0, 1, 2, 3, 4]


def test_fill_patches_roundtrip():
    img = rand_img(3, 3, 16, 16, 6)
    inds = np.array([[0, 3 * 3 * 256 // 3 + 5 * 16 + 2, -1]], np.int64)   # (0,0,0), (1,5,2), invalid
    inds[0, 1] = 1 * 3 * 256 + 5 * 16 + 2
    p = np.full((1, 3, 2, 3, 7, 7), -7., np.float32)
    orc.fill_patches(p, img, inds)
    np.testing.assert_array_equal(p[0, 0], img[0:2, :, 0:7, 0:7])
    np.testing.assert_array_equal(p[0, 1], img[1:3, :, 5:12, 2:9])
    assert np.all(p[0, 2] == -7.)
