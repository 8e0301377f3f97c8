"""Odd shapes through the public API: no NaN/Inf, fast schedule close to parity, oracle agreement on the smallest."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
from oracle import vnlb_oracle as orc
ok = True
for (T, H, W, sigma) in [(2, 7, 7, 20.), (2, 9, 33, 10.), (3, 31, 29, 20.), (5, 57, 101, 30.), (14, 33, 47, 20.), (3, 130, 67, 50.), (2, 64, 64, 5.)]:
    clean = synth.synth_video(T, max(H, 12), max(W, 12))[:, :, :H, :W].copy()
    noisy = synth.add_noise(clean, sigma)
    torch.manual_seed(1)
    dp, bp, _ = vnlb_b200.denoise(noisy, sigma, schedule="parity", verbose=False)
    df, bf, _ = vnlb_b200.denoise(noisy, sigma, schedule="fast", verbose=False)
    fin = bool(torch.isfinite(dp).all() and torch.isfinite(df).all() and torch.isfinite(bp).all())
    pp = vnlb_b200.compute_psnrs(dp, clean).mean(); pf = vnlb_b200.compute_psnrs(df, clean).mean(); pn = vnlb_b200.compute_psnrs(noisy, clean).mean()
    msg = ""
    if T * H * W < 40000:
        torch.manual_seed(1)
        od, ob, _ = orc.denoise(noisy, sigma)
        msg = " max|parity-oracle| %.2e" % np.abs(dp.cpu().numpy() - od).max()
        ok &= np.abs(dp.cpu().numpy() - od).max() < 1e-2
    ok &= fin and abs(pp - pf) < 0.6
    print("%dx%dx%d s%g finite %s psnr noisy %.2f parity %.2f fast %.2f%s" % (W, H, T, sigma, fin, pn, pp, pf, msg), flush=True)
print("ALL OK" if ok else "PROBLEM")
