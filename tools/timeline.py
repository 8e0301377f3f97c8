"""Where does the wall time of one denoise call go?  (CUDA events around coarse regions.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vnlb_b200
from vnlb_b200 import synth, _lib, schedule, proc_nl as pn, color, alloc, impl
T, H, W = 20, 480, 854
noisy = torch.from_numpy(synth.add_noise(synth.synth_video(T, H, W), 20.)).cuda()
marks = []
def mark(name):
    ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append((name, ev, time.time()))
orig_rounds = schedule._rounds_overlapped
def rounds(*a, **k):
    mark("rounds_begin"); r = orig_rounds(*a, **k); mark("rounds_end"); return r
schedule._rounds_overlapped = rounds
orig_finish = schedule.finish_step
def finish(*a, **k):
    r = orig_finish(*a, **k); mark("finish_end"); return r
schedule.finish_step = finish
orig_fast = schedule.proc_nl_fast
def fast(*a, **k):
    mark("step_begin"); return orig_fast(*a, **k)
impl_fast = fast
import vnlb_b200.schedule as S
S.proc_nl_fast = fast
for it in range(3):
    marks.clear()
    torch.cuda.synchronize(); t0 = time.time(); mark("call_begin")
    deno, basic, dt = vnlb_b200.denoise(noisy, 20., verbose=False)
    mark("call_end"); torch.cuda.synchronize()
base = marks[0][1]; bw = marks[0][2]
for name, ev, wall in marks:
    print("%-14s gpu %8.2f ms   host %8.2f ms" % (name, base.elapsed_time(ev), (wall - bw) * 1e3))
print("dtime", dt)
