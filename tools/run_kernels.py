"""Small driver for profiling: runs the search and Bayes kernels on synthetic data.
usage: python tools/run_kernels.py [search|bayes|all] [nrows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vnlb_b200
from vnlb_b200 import synth, search, deno, color, mask as gmask
from vnlb_b200.utils import AttrDict

which = sys.argv[1] if len(sys.argv) > 1 else "all"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = "cuda:0"
T, H, W = 16, 270, 480
clean = synth.synth_video(T, H, W)
noisy = torch.from_numpy(synth.add_noise(clean, 20.)).to(dev)
yuv = color.rgb2yuv(noisy)
# a realistic step-2 input: "basic" = clean + residual noise (sigma 3), not the noisy video itself
yuv_basic = color.rgb2yuv(torch.from_numpy(synth.add_noise(clean, 3.)).to(dev))
params = vnlb_b200.get_params(20.)
for step in (0, 1):
    a = vnlb_b200.get_args(params, 3, step, dev)
    m, ng = gmask.init_mask(noisy.shape, a, dev)
    q = torch.nonzero(m)[::max(1, ng // rows)][:rows].contiguous()
    k = a.npatches
    vals = torch.empty((q.shape[0], k), device=dev)
    inds = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(2):
        ev[0].record()
        if which in ("search", "all"):
            search.exec_sim_search_burst(yuv if step == 0 else yuv_basic, q, vals, inds, None, 20., a)
        ev[1].record()
    if which == "search":
        torch.cuda.synchronize(); print("step", step, "search ms", ev[0].elapsed_time(ev[1]), "queries", q.shape[0]); continue
    if which in ("bayes", "fused"):
        search.exec_sim_search_burst(yuv if step == 0 else yuv_basic, q, vals, inds, None, 20., a)
    if which == "fused":
        images = AttrDict(noisy=yuv, basic=yuv if step == 0 else yuv_basic, deno=torch.zeros_like(yuv), weights=torch.zeros((T, H, W), device=dev))
        for it in range(2):
            ev[2].record()
            deno.bayes_aggregate_fused(images, inds, a)
            ev[3].record()
        torch.cuda.synchronize()
        print("step", step, "fused bayes ms", ev[2].elapsed_time(ev[3]), "rows", q.shape[0])
        continue
    pn = torch.empty((q.shape[0], k, 2, 3, 7, 7), device=dev)
    pb = torch.empty_like(pn)
    search.fill_patches(pn, yuv, inds)
    search.fill_patches(pb, yuv, inds)
    patches = AttrDict(noisy=pn, basic=pb, flat=torch.zeros(q.shape[0], dtype=torch.uint8, device=dev))
    for it in range(2):
        ev[2].record()
        deno.denoise(patches, a, "bayes", inds)
        ev[3].record()
    torch.cuda.synchronize()
    print("step", step, "search ms", ev[0].elapsed_time(ev[1]), "bayes ms", ev[2].elapsed_time(ev[3]), "rows", q.shape[0])
