"""Numerical prototype (float32 numpy) of the production eigen path:
Householder tridiagonalisation -> Sturm bisection (top eigenvalues above the
threshold) -> twisted-factorisation eigenvectors -> back-transformation.
Compares the resulting filter projector with LAPACK's.  Development tool."""
import numpy as np
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
f32 = np.float32

def householder_tridiag(A):
    A = A.astype(f32).copy(); p = A.shape[0]
    d = np.zeros(p, f32); e = np.zeros(p-1, f32); taus = np.zeros(p, f32); V = np.zeros((p, p), f32)
    for k in range(p-2):
        x = A[k+1:, k].copy()
        alpha = x[0]
        sig = f32(np.sum((x[1:]*x[1:]).astype(f32), dtype=f32))
        if sig == 0:
            tau = f32(0); beta = alpha; v = np.zeros_like(x); v[0] = 1
        else:
            beta = f32(-np.copysign(np.sqrt(alpha*alpha + sig, dtype=f32), alpha))
            tau = f32((beta - alpha)/beta)
            v = (x / f32(alpha - beta)).astype(f32); v[0] = 1
        e[k] = beta; d[k] = A[k, k]; taus[k] = tau; V[k+1:, k] = v
        if tau != 0:
            A22 = A[k+1:, k+1:]
            y = (tau * (A22 @ v)).astype(f32)
            a = f32(-0.5) * tau * f32(y @ v)
            w = (y + a * v).astype(f32)
            A[k+1:, k+1:] = (A22 - np.outer(v, w) - np.outer(w, v)).astype(f32)
    d[p-2] = A[p-2, p-2]; d[p-1] = A[p-1, p-1]; e[p-2] = A[p-1, p-2]
    return d, e, taus, V

def sturm_count(d, e2, x, pivmin):
    # number of eigenvalues < x
    cnt = 0; q = f32(1)
    for i in range(len(d)):
        q = f32(d[i] - x) - (f32(e2[i-1] / q) if i > 0 else f32(0))
        if abs(q) < pivmin: q = -pivmin
        cnt += q < 0
    return cnt

def bisect(d, e, idx, lo, hi, iters=30):
    e2 = (e*e).astype(f32); pivmin = f32(max(1e-30, float(e2.max(initial=0))*1e-14))
    lo = f32(lo); hi = f32(hi)
    for _ in range(iters):
        mid = f32(0.5)*(lo+hi)
        if mid <= lo or mid >= hi: break
        if sturm_count(d, e2, mid, pivmin) > idx: hi = mid
        else: lo = mid
    return f32(0.5)*(lo+hi)

def getvec(d, e, lam):
    p = len(d); dp = np.zeros(p, f32); dm = np.zeros(p, f32)
    pivmin = f32(1e-30 + 1e-14*float(np.max(np.abs(d)))) 
    dp[0] = d[0]-lam
    for i in range(p-1):
        if abs(dp[i]) < pivmin: dp[i] = -pivmin
        dp[i+1] = f32(d[i+1]-lam) - f32(e[i]*e[i]/dp[i])
    dm[p-1] = d[p-1]-lam
    for i in range(p-2, -1, -1):
        if abs(dm[i+1]) < pivmin: dm[i+1] = -pivmin
        dm[i] = f32(d[i]-lam) - f32(e[i]*e[i]/dm[i+1])
    gam = dp + dm - (d - lam).astype(f32)
    r = int(np.argmin(np.abs(gam)))
    z = np.zeros(p, f32); z[r] = 1
    for i in range(r-1, -1, -1):
        z[i] = -f32(e[i]/dp[i])*z[i+1]
    for i in range(r, p-1):
        z[i+1] = -f32(e[i]/dm[i+1])*z[i]
    return (z/np.linalg.norm(z)).astype(f32)

def back_transform(Z, taus, V):
    p = Z.shape[0]; Z = Z.copy()
    for k in range(p-3, -1, -1):
        v = V[:, k]; s = (taus[k]*(v @ Z)).astype(f32)
        Z = (Z - np.outer(v, s)).astype(f32)
    return Z

def my_projector(C, sigma2, sigmab2, thresh, rank, refine=False):
    p = C.shape[0]
    d, e, taus, V = householder_tridiag(C)
    tau_eff = f32(thresh*sigma2 + sigmab2)
    e2 = (e*e).astype(f32); pivmin = f32(max(1e-30, float(e2.max(initial=0))*1e-14))
    nless = sturm_count(d, e2, tau_eff, pivmin)
    m = min(p - nless, rank)
    ea = np.abs(np.concatenate([[0], e])) + np.abs(np.concatenate([e, [0]]))
    gmax = f32(np.max(d + ea)); gmax = gmax + f32(1e-6)*abs(gmax)
    lam = np.array([bisect(d, e, p-1-j, tau_eff, gmax) for j in range(m)], f32)
    Z = np.zeros((p, m), f32)
    for j in range(m):
        z = getvec(d, e, lam[j])
        if refine:
            Tz = d*z; Tz[:-1] += e*z[1:]; Tz[1:] += e*z[:-1]
            rq = f32(z @ Tz)
            z = getvec(d, e, rq); 
        Z[:, j] = z
    Vm = back_transform(Z, taus, V)
    ls = lam - np.minimum(lam, f32(sigmab2))
    w = np.where(ls > thresh*sigma2, 1/(1+sigma2/np.maximum(ls, 1e-30)), 0).astype(f32)
    return (Vm*w) @ Vm.T, lam, Vm, w

def ref_projector(C, sigma2, sigmab2, thresh, rank):
    ev, evec = np.linalg.eigh(C.astype(np.float64))
    ev = ev[::-1]; evec = evec[:, ::-1]
    l = ev[:rank]; ls = l - np.minimum(l, sigmab2)
    w = np.where(ls > thresh*sigma2, 1/(1+sigma2/np.maximum(ls, 1e-300)), 0)
    return (evec[:, :rank]*w) @ evec[:, :rank].T, ev

if __name__ == "__main__":
    import inputs as gin
    from oracle import vnlb_oracle as orc
    rng = np.random.RandomState(0)
    worst = 0
    for step in (0, 1):
        for trial in range(6):
            pn, pb, flat = gin.bayes_inputs(step, b=2, seed=100+trial, sigma=20.)
            b, n = pn.shape[:2]
            X = pn.transpose(0, 3, 1, 2, 4, 5).reshape(b, 3, n, -1)
            Bc = pb.transpose(0, 3, 1, 2, 4, 5).reshape(b, 3, n, -1)
            src = X if step == 0 else Bc
            for g in range(b):
                for ch in range(3):
                    Y = src[g, ch] - src[g, ch].mean(0, keepdims=True)
                    C = (Y.T @ Y / n).astype(f32)
                    s2 = 400.; sb2 = 400. if step == 0 else 0.; th = 2.7 if step == 0 else 0.7
                    P, lam, Vm, w = my_projector(C, s2, sb2, th, 39)
                    P2, _, Vm2, _ = my_projector(C, s2, sb2, th, 39, refine=True)
                    Pr, ev = ref_projector(C, s2, sb2, th, 39)
                    Xn = X[g, ch] - X[g, ch].mean(0, keepdims=True)
                    out, out2, outr = Xn @ P, Xn @ P2, Xn @ Pr
                    err = np.linalg.norm(out-outr)/max(np.linalg.norm(outr), 1e-9)
                    err2 = np.linalg.norm(out2-outr)/max(np.linalg.norm(outr), 1e-9)
                    orth = np.abs(Vm.T @ Vm - np.eye(Vm.shape[1])).max() if Vm.shape[1] else 0
                    orth2 = np.abs(Vm2.T @ Vm2 - np.eye(Vm2.shape[1])).max() if Vm2.shape[1] else 0
                    lerr = np.abs(lam - ev[:len(lam)]).max()/ev[0] if len(lam) else 0
                    print(f"step{step+1} t{trial} g{g} c{ch}: m={len(lam):2d} nsel={int((w>0).sum()):2d} relerr(centred)={err:.2e} refined={err2:.2e} orth={orth:.1e}/{orth2:.1e} lam_relerr={lerr:.1e}")
                    worst = max(worst, err2)
    print("worst refined", worst)
