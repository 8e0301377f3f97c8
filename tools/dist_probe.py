"""Where do the intermittent +100..300 ms of a multi-GPU call with gather=True go?  Host wall-clock and device timestamps
around the phases of denoise_distributed.   torchrun --nproc-per-node N tools/dist_probe.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import vnlb_b200
from vnlb_b200 import dist as vdist
from vnlb_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T, H, W = 30, 1080, 1920
clean, flows = synth.synth_video(T, H, W, 123, return_flows=True)
noisy = torch.from_numpy(synth.add_noise(clean, 10., 123)).cuda()
fl = dict(fflow=torch.from_numpy(flows["fflow"]).cuda(), bflow=torch.from_numpy(flows["bflow"]).cuda())
del clean, flows
orig_gather = vdist.gather_bands
log = []


def timed_gather(*a, **k):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig_gather(*a, **k)
    e1.record()
    log.append(("gather", time.perf_counter() - t0, e0, e1))
    return out


vdist.gather_bands = timed_gather
for mode in (True, False, True):
    for it in range(steps if mode else 3):
        log.clear()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = {}
        vdist.denoise_distributed(noisy, 10., flows=fl, stats=st, max_flow=2.0, gather=mode)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        g = ["host %.1f ms dev %.1f ms" % (x[1] * 1e3, x[2].elapsed_time(x[3])) for x in log]
        ms = torch.tensor([dt], device="cuda")
        allms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allms, ms)
        if rank == 0:
            print("gather=%s it %d: per-rank wall ms %s | rank0 gathers: %s | exchange %s | reserved %.1f GB retries %d" % (
                mode, it, [round(float(x), 1) for x in allms], g, [round(x, 1) for x in st.get("exchange_ms", [])],
                torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_stats()["num_alloc_retries"]), flush=True)
dist.destroy_process_group()
