"""Seeded input generators shared by tests/golden/make_golden.py (which runs the
reference on them) and the tests (which run the oracle / the CUDA path on the
same inputs).  numpy RandomState streams are platform independent."""
import numpy as np


def bayes_inputs(step, b=3, seed=7, sigma=20., c=3, ps=7, pt=2):
    """Patch stacks [b,n,pt,c,ps,ps]: low-rank signal (rank ~10) + N(0,sigma^2);
    step 2 adds a basic stack = signal + N(0,3^2)."""
    n = 100 if step == 0 else 60
    p = pt * ps * ps
    rng = np.random.RandomState(seed + step)
    basis = rng.randn(b, c, 10, p).astype(np.float32)
    coef = (rng.randn(b, c, n, 10) * np.linspace(40, 4, 10)).astype(np.float32)
    sig = np.einsum("bcnr,bcrp->bcnp", coef, basis).astype(np.float32) / np.float32(3.)
    sig += (rng.rand(b, c, 1, p) * 200).astype(np.float32)
    if b > 1:
        sig[1] = sig[1] * 0.0 + (rng.rand(c, 1, 1) * 200).astype(np.float32)  # a flat group
    noisy = sig + rng.randn(b, c, n, p).astype(np.float32) * np.float32(sigma)
    basic = sig + rng.randn(b, c, n, p).astype(np.float32) * np.float32(3.)
    if step == 0:
        basic = np.zeros_like(noisy)

    def unflat(x):  # 'b c n (pt ph pw) -> b n pt c ph pw'
        return np.ascontiguousarray(x.reshape(b, c, n, pt, ps, ps).transpose(0, 2, 3, 1, 4, 5))
    flat = np.zeros(b, bool)
    if step == 1 and b > 1:
        flat[1] = True
    return unflat(noisy.astype(np.float32)), unflat(basic.astype(np.float32)), flat


def flat_inputs(seed=11, sigma=20.):
    """6 groups [6,60,2,3,7,7]: noise std sweeps across the flat threshold
    sqrt(gamma)*sigma = 0.447*sigma."""
    rng = np.random.RandomState(seed)
    stds = np.array([0.2, 0.4, 0.44, 0.46, 0.6, 1.0], np.float32) * np.float32(sigma)
    x = rng.randn(6, 60, 2, 3, 7, 7).astype(np.float32) * stds[:, None, None, None, None, None]
    x += (rng.rand(6, 1, 1, 3, 1, 1) * 200).astype(np.float32)
    return x.astype(np.float32)


def agg_inputs(seed=5, T=3, C=3, H=16, W=18, B=4, K=10, ps=7, pt=2):
    rng = np.random.RandomState(seed)
    patches = (rng.rand(B, K, pt, C, ps, ps) * 255).astype(np.float32)
    t = rng.randint(0, T - pt + 1, (B, K))
    y = rng.randint(0, H - ps + 1, (B, K))
    x = rng.randint(0, W - ps + 1, (B, K))
    inds = (t * C * H * W + y * W + x).astype(np.int64)
    inds[2, 3] = -1                      # an invalid row is skipped entirely
    return patches, inds, (T, C, H, W)


def mask_update_inputs(seed=3, T=4, C=3, H=20, W=24, B=6, K=12):
    rng = np.random.RandomState(seed)
    t = rng.randint(0, T, (B, K))
    y = rng.randint(0, H, (B, K))
    x = rng.randint(0, W, (B, K))
    inds = (t * C * H * W + y * W + x).astype(np.int64)
    inds[0, 0] = 0                       # corner: neighbours out of bounds
    inds[0, 1] = (T - 1) * C * H * W + (H - 1) * W + (W - 1)
    inds[4, 5] = -1                      # row 4 dropped
    return inds, (T, C, H, W)


def color_inputs(seed=2):
    rng = np.random.RandomState(seed)
    return (rng.rand(2, 3, 5, 7) * 255).astype(np.float32)


E2E = dict(T=4, H=40, W=48, sigma=20., seed=123, torch_seed=123)
E2E_CASES = dict(
    e2e=E2E,
    e2e_s10=dict(T=4, H=40, W=48, sigma=10., seed=123, torch_seed=123),      # noise levels of BASELINE configs[4]
    e2e_s50=dict(T=4, H=40, W=48, sigma=50., seed=123, torch_seed=123),
    e2e_cfg1=dict(T=3, H=64, W=64, sigma=20., seed=123, torch_seed=123),     # BASELINE configs[0]: davis_64x64-shaped, 3 frames
)
