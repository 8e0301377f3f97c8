"""Does a 854x480x20 call run slower after 1920x1080x30 calls in the same process? (bench.py's config2 extra)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vnlb_b200
from vnlb_b200 import synth


def run(noisy, sigma, flows, n, tag):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        vnlb_b200.denoise(noisy, sigma, verbose=False, flows=flows)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(tag, " ".join("%.1f" % t for t in ts), "reserved GB %.2f" % (torch.cuda.memory_reserved() / 1e9), flush=True)


small = torch.from_numpy(synth.add_noise(synth.synth_video(20, 480, 854), 20.)).cuda()
run(small, 20., None, 4, "small first")
clean, flows = synth.synth_video(30, 1080, 1920, 123, return_flows=True)
big = torch.from_numpy(synth.add_noise(clean, 10., 123)).cuda()
fl = dict(fflow=torch.from_numpy(flows["fflow"]).cuda(), bflow=torch.from_numpy(flows["bflow"]).cuda())
run(big, 10., fl, 2, "big")
run(small, 20., None, 4, "small after big")
torch.cuda.empty_cache()
run(small, 20., None, 4, "small after empty_cache")
del big, fl
torch.cuda.empty_cache()
run(small, 20., None, 4, "small after freeing big")
