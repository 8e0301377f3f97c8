"""PSNR of the throughput schedule vs the reference-exact parity schedule on a mid-size video."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
for (T, H, W, sigma) in [(8, 240, 320, 20.), (8, 240, 320, 10.), (8, 240, 320, 50.)]:
    clean = synth.synth_video(T, H, W); noisy = synth.add_noise(clean, sigma)
    res = {}
    for sched in ("parity", "fast"):
        torch.manual_seed(123); st = {}
        t0 = time.time()
        deno, basic, dt = vnlb_b200.denoise(noisy, sigma, schedule=sched, verbose=False, stats=st)
        res[sched] = (vnlb_b200.compute_psnrs(basic, clean).mean(), vnlb_b200.compute_psnrs(deno, clean).mean(), st["ngroups"], dt, deno.cpu().numpy())
    p, f = res["parity"], res["fast"]
    print("%dx%dx%d sigma %g: parity basic %.3f deno %.3f groups %s %.2fs | fast basic %.3f deno %.3f groups %s %.3fs | delta basic %+.3f deno %+.3f dB, max|fast-parity| %.2f" % (
        W, H, T, sigma, p[0], p[1], p[2], p[3], f[0], f[1], f[2], f[3], f[0] - p[0], f[1] - p[1], np.abs(p[4] - f[4]).max()))
