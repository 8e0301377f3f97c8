"""Round-2 study (CPU, numpy): can the per-group eigen-stage be done with GEMM-shaped work?

Block subspace iteration V <- orth(C V) with a block of b = 48 > rank = 39 columns, products rounded to
TF32 (10-bit mantissa, emulated) during the iterations, followed by ONE Rayleigh-Ritz step in FP32.
Measures, on the covariance matrices of a real oracle run, how many iterations are needed until the filter
projector  P = V_m diag(w) V_m^T  matches the exact one (float64 eigh) to the parity tolerance."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
from oracle import vnlb_oracle as orc

f32 = np.float32

def tf32(x):
    """round-to-nearest to a 10-bit mantissa (TF32 operand precision)"""
    xi = np.ascontiguousarray(x, dtype=f32).view(np.uint32)
    r = ((xi + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).astype(np.uint32)
    return r.view(f32)

def projector(evals, evecs, s2, sb2, th, rank):
    l = evals[:rank]; ls = l - np.minimum(l, sb2)
    w = np.where(ls > th * s2, 1.0 / (1.0 + s2 / np.maximum(ls, 1e-30)), 0.0)
    V = evecs[:, :rank]
    return (V * w) @ V.T, int((w > 0).sum())

def subspace(C, b, iters, lowprec, rng):
    p = C.shape[0]
    V = np.linalg.qr(rng.standard_normal((p, b)))[0].astype(f32)
    Cl = tf32(C) if lowprec else C
    for _ in range(iters):
        W = (Cl.astype(f32) @ (tf32(V) if lowprec else V)).astype(f32)
        G = (W.T @ W).astype(f32)                               # Cholesky-QR (twice for stability)
        L = np.linalg.cholesky(G.astype(np.float64) + 1e-30 * np.eye(b)).astype(f32)
        V = np.linalg.solve(L.astype(np.float64), W.T.astype(np.float64)).T.astype(f32)
        G = (V.T @ V).astype(f32)
        L = np.linalg.cholesky(G.astype(np.float64)).astype(f32)
        V = np.linalg.solve(L.astype(np.float64), V.T.astype(np.float64)).T.astype(f32)
    H = (V.T @ (C.astype(f32) @ V)).astype(f32)                 # Rayleigh-Ritz in FP32
    th_, Y = np.linalg.eigh(H.astype(np.float64))
    return th_[::-1].astype(f32), (V @ Y[:, ::-1]).astype(f32)

def collect_matrices(step, ngroups=40):
    T, H, W, sigma = 6, 72, 96, 20.
    clean = orc.synth_video(T, H, W, 3); noisy = orc.add_noise(clean, sigma, 3)
    a = orc.get_args(orc.default_params(sigma), 3, step)
    yn = orc.rgb2yuv(noisy)
    yb = orc.rgb2yuv((clean + np.random.RandomState(1).randn(*clean.shape) * 4).astype(f32))   # stand-in for `basic`
    rs = np.random.RandomState(0)
    q = np.stack([rs.randint(0, T - 1, ngroups), rs.randint(0, H - 6, ngroups), rs.randint(0, W - 6, ngroups)], 1).astype(np.int64)
    k = a.npatches
    vals = np.full((ngroups, k), np.inf, f32); inds = np.full((ngroups, k), -1, np.int64)
    srch = yn if step == 0 else yb
    orc.exec_sim_search_burst(srch, q, vals, inds, None, sigma, a)
    pn = np.zeros((ngroups, k, 2, 3, 7, 7), f32); orc.fill_patches(pn, srch, inds)
    X = pn.transpose(0, 3, 1, 2, 4, 5).reshape(ngroups, 3, k, -1)
    X = X - X.mean(2, keepdims=True)
    return [(X[g, c].T @ X[g, c] / k).astype(f32) for g in range(ngroups) for c in range(3)], a

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for step in (0, 1):
        mats, a = collect_matrices(step)
        s2, sb2, th = a.sigma2, a.sigmab2, a.thresh
        print("step %d: %d matrices, tau_eff = %.0f" % (step + 1, len(mats), th * s2 + sb2))
        for lowprec in (False, True):
            for iters in (2, 4, 6, 8, 12):
                errs, mism, ms = [], 0, []
                for C in mats:
                    ev, evec = np.linalg.eigh(C.astype(np.float64)); ev = ev[::-1]; evec = evec[:, ::-1]
                    Pref, m = projector(ev, evec, s2, sb2, th, 39)
                    thv, V = subspace(C, 48, iters, lowprec, rng)
                    P, m2 = projector(thv.astype(np.float64), V.astype(np.float64), s2, sb2, th, 39)
                    mism += (m != m2); ms.append(m)
                    nr = max(np.linalg.norm(Pref), 1e-12)
                    errs.append(np.linalg.norm(P - Pref) / nr if m > 0 else np.linalg.norm(P))
                print("  %-5s iters %2d: count mismatches %3d/%d   projector rel err median %.1e  max %.1e   (m: median %d, max %d)" % (
                    "tf32" if lowprec else "fp32", iters, mism, len(mats), np.median(errs), np.max(errs), np.median(ms), np.max(ms)))
