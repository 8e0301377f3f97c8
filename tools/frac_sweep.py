import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
T, H, W = 20, 480, 854
clean = synth.synth_video(T, H, W); noisy = torch.from_numpy(synth.add_noise(clean, 20.)).cuda()
for frac, qmin, cap in [(1/8, 4096, 16384), (1/12, 4096, 16384), (1/16, 4096, 16384), (1/16, 2048, 16384), (1/24, 2048, 16384), (1/32, 2048, 16384), (1/12, 4096, 8192), (1/16, 2048, 8192)]:
    params = vnlb_b200.get_params(20.)
    params["fast_frac"] = [frac, frac]; params["fast_min"] = [qmin, qmin]; params["fast_cap"] = [cap, cap]
    for it in range(3):
        st = {}
        torch.cuda.synchronize(); t0 = time.time()
        deno, basic, dt = vnlb_b200.denoise(noisy, 20., verbose=False, params=params, stats=st)
        torch.cuda.synchronize(); el = time.time() - t0
    ps = vnlb_b200.compute_psnrs(deno, clean).mean(); pb = vnlb_b200.compute_psnrs(basic, clean).mean()
    print("cap %d frac 1/%d min %d: %.1f ms groups %s rounds %s psnr basic %.3f deno %.3f" % (cap, round(1/frac), qmin, el*1e3, st["ngroups"], st["nrounds"], pb, ps))
