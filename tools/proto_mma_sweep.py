"""Round-2 study (CPU): accuracy of the Householder tridiagonalisation when the rank-2 update
B <- B - v w^T - w v^T runs on the tensor cores as ONE 3xTF32 product per tile,
    [v_hi v_hi v_lo w_hi w_hi w_lo] . [w_hi w_lo w_hi v_hi v_lo v_hi]^T      (K = 6 of the 8 slots of mma.m16n8k8),
with FP32 accumulation (the matrix stays FP32 in the accumulator fragments).  The emulation rounds the operand
splits to TF32 (10-bit mantissa, round to nearest) and drops the lo*lo terms, exactly what the product would do.

Reports, over random covariance matrices of the step-1 shape (n = 100 patches, p = 98), the error of the
eigenvalues and of the Wiener-filter projector against float64 for (a) plain FP32 updates, (b) 3xTF32 updates,
(c) single-pass TF32 updates.  usage: python tools/proto_mma_sweep.py [trials]"""
import sys
import numpy as np


def tf32(x):
    """round-to-nearest TF32 (keeps 10 explicit mantissa bits) of a float32 array"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000 + ((u >> 13) & 1) - 1) & 0xFFFFE000      # round half to even on bit 13
    return (u & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def rank2(B, v, w, mode):
    v = v.astype(np.float32); w = w.astype(np.float32)
    if mode == "fp32":
        return (B - np.outer(v, w) - np.outer(w, v)).astype(np.float32)
    vh, wh = tf32(v), tf32(w)
    if mode == "tf32":
        return (B - np.outer(vh, wh) - np.outer(wh, vh)).astype(np.float32)
    vl, wl = tf32(v - vh), tf32(w - wh)
    upd = (np.outer(vh, wh) + np.outer(vh, wl) + np.outer(vl, wh) +
           np.outer(wh, vh) + np.outer(wh, vl) + np.outer(wl, vh))
    return (B - upd.astype(np.float32)).astype(np.float32)


def tridiag(A, mode):
    """forward Householder tridiagonalisation in float32 with the chosen rank-2 update; returns d, e, Q"""
    B = A.astype(np.float32).copy()
    q = B.shape[0]
    Q = np.eye(q, dtype=np.float64)
    for k in range(q - 2):
        x = B[k + 1:, k].astype(np.float32)
        alpha = x[0]
        nrm = np.float32(np.sqrt(np.dot(x.astype(np.float64), x.astype(np.float64))))
        if nrm == abs(alpha):
            continue
        beta = np.float32(-np.copysign(nrm, alpha))
        tau = np.float32((beta - alpha) / beta)
        v = (x / np.float32(alpha - beta)).astype(np.float32); v[0] = 1.0
        T = B[k + 1:, k + 1:]
        pvec = (tau * (T @ v)).astype(np.float32)
        w = (pvec - np.float32(0.5) * tau * np.dot(pvec, v) * v).astype(np.float32)
        B[k + 1:, k + 1:] = rank2(T, v, w, mode)
        B[k + 1, k] = B[k, k + 1] = beta
        B[k + 2:, k] = 0; B[k, k + 2:] = 0
        vv = np.zeros(q); vv[k + 1:] = v
        Q = Q - float(tau) * np.outer(Q @ vv, vv)
    return np.diag(B).astype(np.float64), np.diag(B, 1).astype(np.float64), Q


def projector(lam, V, sigma2, thresh, rank):
    order = np.argsort(-lam)[:rank]
    P = np.zeros((V.shape[0], V.shape[0]))
    for r in order:
        if lam[r] > thresh * sigma2:
            P += (1.0 / (1.0 + sigma2 / lam[r])) * np.outer(V[:, r], V[:, r])
    return P


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rs = np.random.RandomState(0)
    sigma2, thresh, rank = 400.0, 2.7, 39
    out = {m: [] for m in ("fp32", "3xtf32", "tf32")}
    for t in range(trials):
        n, p = 100, 98
        nsig = rs.randint(0, 12)                      # a few strong components + noise, like a real group
        Y = rs.randn(n, p) * 20.0
        if nsig:
            Y += (rs.randn(n, nsig) * rs.uniform(20, 200, nsig)) @ rs.randn(nsig, p) / np.sqrt(p) * 3
        Y -= Y.mean(0)
        C = (Y.T @ Y / n)
        lam64, V64 = np.linalg.eigh(C)
        P64 = projector(lam64, V64, sigma2, thresh, rank)
        for mode in out:
            d, e, Q = tridiag(C, mode)
            T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
            lam, Z = np.linalg.eigh(T)
            V = Q @ Z
            P = projector(lam, V, sigma2, thresh, rank)
            out[mode].append((np.abs(lam - lam64).max() / lam64.max(), np.linalg.norm(P - P64) / max(np.linalg.norm(P64), 1e-30)))
    for mode, v in out.items():
        v = np.array(v)
        print("%-7s eigenvalue error / lambda_max: median %.2e max %.2e | filter projector (relative, Frobenius): median %.2e max %.2e"
              % (mode, np.median(v[:, 0]), v[:, 0].max(), np.median(v[:, 1]), v[:, 1].max()))


if __name__ == "__main__":
    main()
