"""BASELINE config 5 on one GPU: 1920x1080x30 synthetic RGB, sigma 10 / 20 / 50, zero flow and a synthetic flow field."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
T, H, W = 30, 1080, 1920
clean = synth.synth_video(T, H, W)
yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
ff = np.zeros((T, 2, H, W), np.float32)
ff[:, 0] = 1.3 * np.sin(yy / 97.)[None] + 0.5      # smooth horizontal drift (px/frame)
ff[:, 1] = 0.9 * np.cos(xx / 131.)[None]
flows = dict(fflow=torch.from_numpy(ff).cuda(), bflow=torch.from_numpy(-ff).cuda())
out = {}
for sigma in (10., 20., 50.):
    noisy = torch.from_numpy(synth.add_noise(clean, sigma)).cuda()
    for name, fl in (("noflow", None), ("flow", flows)):
        for it in range(2):
            st = {}
            torch.cuda.synchronize(); t0 = time.time()
            deno, basic, dt = vnlb_b200.denoise(noisy, sigma, flows=fl, verbose=False, stats=st)
            torch.cuda.synchronize(); el = time.time() - t0
        r = dict(seconds=el, Mpx_per_s=T * H * W / 1e6 / el, groups=st["ngroups"], rounds=st["nrounds"],
                 psnr_noisy=float(vnlb_b200.compute_psnrs(noisy, clean).mean()),
                 psnr_basic=float(vnlb_b200.compute_psnrs(basic, clean).mean()),
                 psnr_deno=float(vnlb_b200.compute_psnrs(deno, clean).mean()),
                 max_mem_GB=torch.cuda.max_memory_allocated() / 1e9)
        out["sigma%d_%s" % (sigma, name)] = r
        print(sigma, name, json.dumps(r), flush=True)
json.dump(out, open("gpurun_out/config5_r1b.json", "w"), indent=1)
