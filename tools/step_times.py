import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth
noisy = torch.from_numpy(synth.add_noise(synth.synth_video(20, 480, 854), 20.)).cuda()
for mode in ("async", True, False):
    params = vnlb_b200.get_params(20.); params["fast_overlap"] = [mode, mode]
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vnlb_b200.denoise(noisy, 20., verbose=False, params=params)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(mode, " ".join("%.0f" % t for t in ts), flush=True)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["allocation.all.allocated"], torch.cuda.memory_stats().get("num_device_alloc"))
import bench, gc
for label, use_sampler, pin in (("sampler", True, False), ("pinned+result", False, True), ("gc.disable", False, True)):
    s = bench.ClockSampler(0)
    if use_sampler: s.start()
    if label == "gc.disable": gc.disable()
    params = vnlb_b200.get_params(20.)
    keep = {}
    pinned = torch.from_numpy(synth.add_noise(synth.synth_video(20, 480, 854), 20.)).pin_memory() if pin else None
    outp = torch.empty_like(pinned).pin_memory() if pin else None
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        keep["d"], keep["b"], _ = vnlb_b200.denoise(pinned if pin else noisy, 20., verbose=False, params=params)
        if pin: outp.copy_(keep["d"], non_blocking=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    if use_sampler: s.stop()
    print(label, " ".join("%.0f" % t for t in ts), flush=True)
