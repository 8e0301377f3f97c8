"""Eager two-stream rounds vs CUDA-graph replays (fast_graph) end to end: python tools/graph_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vnlb_b200
from vnlb_b200 import _lib as L
from vnlb_b200 import synth

for (T, H, W) in ((3, 64, 64), (8, 128, 160), (10, 240, 320), (20, 480, 854)):
    noisy = torch.from_numpy(synth.add_noise(synth.synth_video(T, H, W), 20.)).cuda()
    for graph in (False, True):
        params = vnlb_b200.get_params(20.)
        params["fast_graph"] = [graph, graph]
        ts = []
        for it in range(6):
            st = {}
            n0 = int(L.lib.vnlb_kernel_launches())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            vnlb_b200.denoise(noisy, 20., verbose=False, stats=st, params=params)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("%dx%dx%d" % (W, H, T), "graph" if graph else "eager", " ".join("%.2f" % t for t in ts), "groups", st.get("ngroups"),
              "rounds", st.get("nrounds"), "kernels enqueued by the library in the last call", int(L.lib.vnlb_kernel_launches()) - n0, flush=True)
