#!/usr/bin/env python
"""bench.py -- vnlb.denoise throughput (steps 1+2) in Mpx/s on synthetic video.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

Workload: BASELINE.json configs[1] -- 854x480 DAVIS-shaped synthetic RGB, 20
frames, sigma=20, no flow -- on one GPU.  With N GPUs the video is N bands of
480 rows stacked vertically (854 x 480N x 20; N=8 is 65.6 Mpx, the size of
configs[4]); each rank owns one band of reference pixels, so per-GPU work is
fixed ("weak" scaling) and the accumulators are summed with one all-reduce per
step.  A "step" of this bench = one full vnlb.denoise call (VNLB steps 1+2).

One JSON line is printed by rank 0 (see README / DESIGN.md for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vnlb.denoise Mpx/s (steps 1+2)"
SIGMA = 20.0
BASE = dict(T=20, H=480, W=854)
CPU_SAMPLE = dict(T=6, H=96, W=128)
# multi-GPU band partition (vnlb_b200/dist.py): "auto" (default), "snake", "plain" or "weighted"; override for experiments only
BALANCE = {"plain": False, "snake": "snake", "weighted": "weighted", "auto": "auto"}[os.environ.get("VNLB_BALANCE", "auto")]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read through NVML from the benchmark
    thread itself right after each step is enqueued/finished (a few microseconds per query).  A background
    poller (nvidia-smi -lms or an NVML thread) was measured to cause sporadic +50..100 ms stalls of single
    steps on this box, so no second thread is used."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.ok, self.max_mhz = index, [], False, 0.0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.sample()
            self.rows.clear()
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.rows.append((mhz, mask))
        except Exception:
            pass

    def summary(self):
        if not self.ok or not self.rows:
            return None
        reasons = sorted({name for _, m in self.rows for bit, name in self.REASONS.items() if m & bit})
        return dict(sm_mhz=float(np.median([r[0] for r in self.rows])), sm_max_mhz=self.max_mhz, reasons=reasons,
                    samples=len(self.rows), source="nvml, sampled while the timed steps execute")


def make_video(n_gpus):
    from vnlb_b200 import synth
    T, H, W = BASE["T"], BASE["H"] * n_gpus, BASE["W"]
    clean = synth.synth_video(T, H, W, 123)
    return clean, synth.add_noise(clean, SIGMA, 123)


def cpu_reference_run(threads=None):
    """The reference's CPU implementation of the path (oracle port: the
    reference's own stages restated in numpy + the C/OpenMP restatement of the
    absent vpss search) on a bounded sample of the workload.  Returns
    (Mpx/s, seconds, cores, sample description)."""
    from oracle import vnlb_oracle as orc
    import torch
    clean, noisy = make_video(1)
    s = CPU_SAMPLE
    crop = np.ascontiguousarray(noisy[:s["T"], :, :s["H"], :s["W"]])
    torch.manual_seed(123)
    t0 = time.time()
    orc.denoise(crop, SIGMA)
    dt = time.time() - t0
    px = s["T"] * s["H"] * s["W"]
    cores = max(orc.lib().oracle_num_threads(), torch.get_num_threads())
    desc = "top-left %dx%dx%d crop of the workload, full vnlb.denoise (steps 1+2), default_params" % (
        s["W"], s["H"], s["T"])
    return px / 1e6 / dt, dt, cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, cores, desc = cpu_reference_run()
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    line = dict(metric=METRIC, value=v, unit="Mpx/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="854x480x20 synthetic RGB, sigma=20, no flow (BASELINE configs[1]); "
                                     "each step = bounded sample: " + desc),
                cpu_baseline=dict(value=v, unit="Mpx/s", cores=cores, kind="port", sample=desc),
                e2e=dict(value=v, unit="Mpx/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import vnlb_b200
    from vnlb_b200 import _lib, dist as vdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    assert world == n_gpus or world == 1 and n_gpus == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    clean, noisy = make_video(n_gpus)
    T, C, H, W = noisy.shape
    mpx = T * H * W / 1e6
    noisy_pin = torch.from_numpy(noisy).pin_memory()
    noisy_dev = noisy_pin.to(device)
    out_pin = torch.empty_like(noisy_pin).pin_memory()
    params = vnlb_b200.get_params(SIGMA)

    def call(x, stats=None):
        if world > 1:
            return vdist.denoise_distributed(x, SIGMA, schedule="fast", params=params, stats=stats, device=device,
                                             balance=BALANCE)
        return vnlb_b200.denoise(x, SIGMA, gpuid=local_rank, verbose=False, schedule="fast", params=params, stats=stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step = {}

    def timed(fn, steps, tag):
        import gc
        gc.collect()
        gc.disable()        # a full collection of the interpreter heap (~100 ms) inside a step is a host stall, not GPU time
        try:
            return _timed(fn, steps, tag)
        finally:
            gc.enable()

    def _timed(fn, steps, tag):
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        barrier()
        per_step[tag] = [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(steps)]
        ms = torch.tensor([evs[0].elapsed_time(evs[steps])], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # the clock sampler starts before the warm-up so that spawning nvidia-smi never lands in a timed region;
    # only the samples taken during the timed regions are reported
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    result = {}

    def step_dev():
        result["deno"], result["basic"], _ = call(noisy_dev)

    # ---- warm-up: the same step as the timed one, so kernels are loaded and the caching allocator has
    #      reached its steady state (a first-time cudaMalloc inside a step costs 50-100 ms) ----
    for _ in range(args.warmup):
        step_dev()

    # ---- device-resident throughput (`value`) ----
    l0 = int(_lib.lib.vnlb_kernel_launches())   # kernels launched by libvnlb_b200.so, counted in the library

    if rank == 0:   # sample the clocks every 64th kernel launch of the timed steps: the GPU is busy at those moments
        _lib.on_launch = lambda n: sampler.sample() if n % 64 == 0 else None
    ms = timed(step_dev, args.steps, "value")
    launches = (int(_lib.lib.vnlb_kernel_launches()) - l0) // max(args.steps, 1)

    # ---- end to end through the public API with host buffers (`e2e`) ----
    def step_e2e():
        d, _, _ = call(noisy_pin)
        if rank == 0:
            out_pin.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, "e2e")
    _lib.on_launch = None
    clocks = sampler.summary() if rank == 0 else None

    # ---- per-stage device time -> roofline of the dominant kernel (rank 0 timing, max over ranks skipped) ----
    _lib.timer = _lib.StageTimer()
    st2 = {}
    call(noisy_dev, st2)
    stage = _lib.timer.summary()
    _lib.timer = None
    ngroups = [int(g) for g in st2.get("ngroups", [0, 0])]
    per_rank = None
    if world > 1:   # load balance report: groups and device time of the stage pass on every rank
        mine = torch.tensor([float(sum(ngroups)), float(sum(v["ms"] for v in stage.values()))], device=device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        mine_steps = torch.tensor(per_step["value"], device=device, dtype=torch.float32)
        all_steps = [torch.zeros_like(mine_steps) for _ in range(world)]
        dist.all_gather(all_steps, mine_steps)
        per_rank = dict(groups=[int(a[0].item()) for a in allr], stage_ms=[round(float(a[1].item()), 1) for a in allr],
                        value_step_ms=[[round(float(x), 1) for x in a.tolist()] for a in all_steps],
                        allreduce_ms_rank0=[round(x, 2) for x in st2.get("allreduce_ms", [])])

    if rank == 0:
        hbm_peak, peak_kind = load_peaks()
        deno = result["deno"].cpu().numpy()
        basic = result["basic"].cpu().numpy()
        psnr = dict(noisy=float(vnlb_b200.compute_psnrs(noisy, clean).mean()),
                    basic=float(vnlb_b200.compute_psnrs(basic, clean).mean()),
                    deno=float(vnlb_b200.compute_psnrs(deno, clean).mean()))
        # stage keys are "<stage>_s<step>"; merge per stage for the summary
        merged = {}
        for k, v in stage.items():
            base = k.rsplit("_s", 1)[0]
            m = merged.setdefault(base, dict(ms=0.0, launches=0))
            m["ms"] += v["ms"]
            m["launches"] += v["launches"]
        step_ms = sum(v["ms"] for v in merged.values())
        dom = max(merged, key=lambda k: merged[k]["ms"]) if merged else None
        # algorithmic work per group (SURVEY 8d / DESIGN.md section 5), per VNLB step: k = 100 / 60, p = 98, C = 3
        #   fused Bayes: gather k*D*4 (noisy; + basic in step 2) + scatter k*(D+p)*4 ; nominal flop count 35.8 / 31.7 MFLOP
        #   search: N_cand*p*C_d*3 flops, (k*12 + 24) bytes
        work = {
            "bayes": [dict(bytes=100 * 294 * 4 + 100 * 392 * 4, flops=35.8e6),
                      dict(bytes=2 * 60 * 294 * 4 + 60 * 392 * 4, flops=31.7e6)],
            "search": [dict(bytes=100 * 12 + 24, flops=9477 * 98 * 1 * 3), dict(bytes=60 * 12 + 24, flops=9477 * 98 * 3 * 3)],
            "mask_fill_flat": [dict(bytes=100 * 294 * 4, flops=0.), dict(bytes=2 * 60 * 294 * 4, flops=0.)],
            "aggregate": [dict(bytes=100 * 294 * 4, flops=100 * 392.), dict(bytes=60 * 294 * 4, flops=60 * 392.)],
        }
        # DRAM bytes per group of the fused Bayes stage from the ncu --set full capture (profiles/r1b_summary.md, all kernels
        # of a call summed: the per-problem workspace the split kernels hand over is what reaches DRAM; the gathers hit L2)
        ncu_dram_bytes_per_group = {"bayes": [260e3, 104e3]}
        roof = None
        if dom:
            d = merged[dom]
            alg_bytes = sum(work[dom][s]["bytes"] * ngroups[s] for s in (0, 1))
            alg_flops = sum(work[dom][s]["flops"] * ngroups[s] for s in (0, 1))
            sec = d["ms"] / 1e3
            traffic = None
            if dom in ncu_dram_bytes_per_group:
                traffic = sum(ncu_dram_bytes_per_group[dom][s] * ngroups[s] for s in (0, 1)) / max(d["launches"], 1)
            roof = dict(kernel="vnlb_bayes_aggregate_fused: cov_tridiag_kernel + tridiag_tail_kernel + bayes_kernel<fused,split> "
                               "(step 1), gram_tridiag_kernel + tridiag_tail_kernel + bayes_kernel<fused,gram,split> (step 2)" if dom == "bayes" else dom, bound="hbm",
                        achieved=alg_bytes / sec / 1e9, peak=hbm_peak, unit="GB/s",
                        frac=alg_bytes / sec / 1e9 / hbm_peak, traffic=traffic, peak_kind=peak_kind,
                        algorithmic_bytes_per_launch=alg_bytes / max(d["launches"], 1),
                        launches=d["launches"], avg_launch_ms=d["ms"] / max(d["launches"], 1),
                        launch_unit="one C-ABI call of the fused Bayes stage = one round of groups (4 kernels in step 1, 3 in step 2)",
                        share_of_step=d["ms"] / max(step_ms, 1e-9),
                        note="not an HBM-bound stage (arithmetic intensity ~130 flop/B, ridge ~11; ncu: DRAM 1-7 % of peak): its kernels "
                             "are bound by the shared-memory data pipe (58-77 % busy) and dependent chains; the HBM fraction is small "
                             "by construction, see fp32 and profiles/r1b_summary.md",
                        fp32=dict(achieved_tflops=alg_flops / sec / 1e12, nominal_peak_tflops=74.4,
                                  frac=alg_flops / sec / 1e12 / 74.4,
                                  note="nominal LAPACK-style flop count of SURVEY 8d over nominal 148 SM x 128 lanes x 2 x 1.965 GHz"))
        stage = {k: v for k, v in sorted(stage.items())}
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            v, dt, cores, desc = cpu_reference_run()
            cpu = dict(value=v, unit="Mpx/s", cores=cores, kind="port", sample=desc, seconds=dt)
        line = dict(
            metric=METRIC, value=mpx * args.steps / (ms / 1e3), unit="Mpx/s", n_gpus=n_gpus, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload="%dx%dx%d synthetic RGB, sigma=20, no flow (BASELINE configs[1]%s)" % (
                W, H, T, "" if n_gpus == 1 else ", %d bands of 480 rows, one per GPU" % n_gpus),
                params="default_params: 7x7x2 patches, 27x27 window, +-6 frames, k=100/60, rank 39",
                schedule="fast", l2="working set (noisy+basic+accumulators %.0f MB) exceeds the 126 MB L2" % (
                    T * H * W * 4 * 10 / 1e6)),
            e2e=dict(value=mpx * args.steps / (ms_e2e / 1e3), unit="Mpx/s", ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=int(noisy.nbytes) * max(world, 1), d2h_bytes_per_step=int(noisy.nbytes)),
            gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu,
            stages_ms={k: round(v["ms"], 3) for k, v in stage.items()}, groups_per_step=ngroups, psnr=psnr,
            rounds=st2.get("nrounds"), per_rank=per_rank, per_step_ms_rank0=per_step)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames", type=int, default=None, help="override frame count (debugging only)")
    args = ap.parse_args()
    if args.frames:
        BASE["T"] = args.frames
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
