import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
T, H, W, sigma = 10, 480, 854, 20.
clean = synth.synth_video(T, H, W); noisy = torch.from_numpy(synth.add_noise(clean, sigma)).cuda()
def run(sched, frac=None, qmin=None):
    params = vnlb_b200.get_params(sigma)
    if frac: params["fast_frac"] = [frac, frac]
    if qmin: params["fast_min"] = [qmin, qmin]
    torch.manual_seed(1); st = {}
    torch.cuda.synchronize(); t0 = time.time()
    deno, basic, dt = vnlb_b200.denoise(noisy, sigma, schedule=sched, verbose=False, params=params, stats=st)
    torch.cuda.synchronize()
    print(sched, frac, qmin, "groups", st["ngroups"], "rounds", st.get("nrounds"), "%.3fs" % (time.time() - t0),
          "psnr %.3f %.3f" % (vnlb_b200.compute_psnrs(basic, clean).mean(), vnlb_b200.compute_psnrs(deno, clean).mean()), flush=True)
run("fast", 1/8, 4096); run("fast", 1/8, 4096); run("fast", 1/32, 2048); run("fast", 1/128, 512); run("fast", 1/512, 128)
run("parity")
