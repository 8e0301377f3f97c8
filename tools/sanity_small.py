"""Small fused-Bayes run for compute-sanitizer: python tools/sanity_small.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vnlb_b200
from vnlb_b200 import synth, search, deno, color, mask as gmask
from vnlb_b200.utils import AttrDict
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dev = "cuda:0"
T, H, W = 6, 64, 96
clean = synth.synth_video(T, H, W)
noisy = torch.from_numpy(synth.add_noise(clean, 20.)).to(dev)
yuv = color.rgb2yuv(noisy)
yb = color.rgb2yuv(torch.from_numpy(synth.add_noise(clean, 3.)).to(dev))
params = vnlb_b200.get_params(20.)
for step in (0, 1):
    a = vnlb_b200.get_args(params, 3, step, dev)
    m, ng = gmask.init_mask(noisy.shape, a, dev)
    q = torch.nonzero(m)[::max(1, ng // rows)][:rows].contiguous()
    k = a.npatches
    vals = torch.empty((q.shape[0], k), device=dev)
    inds = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    search.exec_sim_search_burst(yuv if step == 0 else yb, q, vals, inds, None, 20., a)
    inds[3, 5] = -1       # an invalid row
    images = AttrDict(noisy=yuv, basic=yuv if step == 0 else yb, deno=torch.zeros_like(yuv), weights=torch.zeros((T, H, W), device=dev))
    deno.bayes_aggregate_fused(images, inds, a)
    torch.cuda.synchronize()
    print("step", step, "ok", float(images.deno.abs().sum()), float(images.weights.sum()))
