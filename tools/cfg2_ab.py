"""854x480x20 (BASELINE configs[1]) end to end with the two search kernels: python tools/cfg2_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vnlb_b200
from vnlb_b200 import _lib as L
from vnlb_b200 import synth

noisy = torch.from_numpy(synth.add_noise(synth.synth_video(20, 480, 854), 20.)).cuda()
for path in (1, 0, 1, 0):
    L.lib.vnlb_set_search_path(path)
    ts = []
    for it in range(5):
        st = {}
        L.timer = L.StageTimer() if it == 4 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        vnlb_b200.denoise(noisy, 20., verbose=False, stats=st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    summ = {k: round(v["ms"], 1) for k, v in L.timer.summary().items()}
    L.timer = None
    print("path", path, " ".join("%.1f" % t for t in ts), st.get("ngroups"), st.get("nrounds"), summ, flush=True)
