#!/usr/bin/env python
"""bench.py -- vnlb.denoise throughput (steps 1+2) in Mpx/s on synthetic video.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port, all host cores)

Workload (every N): BASELINE.json configs[4] -- ONE fixed 1920x1080 synthetic RGB video of 30 frames with the
generator's analytic ("precomputed") flows, sigma = 10 as `value` and sigma = 50 as `sigma50`; with N GPUs the SAME
video is split into N row bands with halos (vnlb_b200/dist.py), i.e. strong scaling.  A "step" of this bench = one full
vnlb.denoise call (VNLB steps 1+2).  At N = 1 the line also carries the other BASELINE configurations measured in the
same run (configs[1] 854x480x20, configs[2] search microbench, configs[3] Bayes microbench, configs[0] 64x64x3 on the
CPU) and the PSNR of the benched (throughput) schedule against the reference-exact schedule and the CPU oracle.

One JSON line is printed by rank 0 (see README / DESIGN.md section 7 for the keys).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vnlb.denoise Mpx/s (steps 1+2)"
WORK = dict(T=30, H=1080, W=1920)           # BASELINE configs[4]
SIGMA, SIGMA_ALT = 10.0, 50.0
MAX_FLOW = 2.0                              # the generator's objects move at most 2 px / frame (vnlb_b200/synth.py)
CFG2 = dict(T=20, H=480, W=854, sigma=20.0)   # BASELINE configs[1]
CPU_SAMPLE = dict(T=6, H=96, W=128)         # bounded sample of the workload for the CPU arm (top-left crop, with flows)
PARITY_SUB = dict(T=10, H=480, W=854)       # sub-video on which the benched schedule is compared with the parity schedule
WORKLOAD_DESC = ("1920x1080x30 synthetic RGB, sigma=10 (sigma=50 in `sigma50`), precomputed (analytic) flows "
                 "(BASELINE configs[4]); one fixed video split over the GPUs in row bands + halos")


def load_peaks():
    """(HBM GB/s, kind, FP32 TFLOP/s, kind, file record).  HBM: driver-written MEASURED_PEAKS.json.  FP32 CUDA-core peak:
    profiles/fp32_peak.json, written by tools/measure_fp32_peak.py (tools/fp32_peak.cu run on a B200); nominal otherwise."""
    hbm, hbm_kind = 6650.0, "fallback"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm, hbm_kind = float(json.load(open(p)).get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    fp32, fp32_kind, rec = 74.4, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz", None
    p = os.path.join(ROOT, "profiles", "fp32_peak.json")
    if os.path.exists(p):
        rec = json.load(open(p))
        fp32, fp32_kind = float(rec["fp32_tflops"]), "measured (profiles/fp32_peak.json, tools/fp32_peak.cu)"
    return hbm, hbm_kind, fp32, fp32_kind, rec


def load_ncu_traffic():
    """DRAM bytes per group of the fused Bayes stage (dram__bytes_read.sum + dram__bytes_write.sum of its kernels from one
    `ncu --set full` capture, divided by the groups of the captured launch), read from the committed summary."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read through NVML from the benchmark thread itself right
    after kernel launches (a few microseconds per query).  A background poller was measured to cause sporadic
    +50..100 ms stalls of single steps, so no second thread is used."""
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


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(threads):
    """The reference's CPU implementation of the path -- the oracle port: the reference's own stages restated in numpy
    (per-matrix LAPACK eigh spread over `threads` host threads) + the C/OpenMP restatement of the absent vpss search --
    on a bounded sample of the workload.  Never imports vnlb_b200.  Returns (Mpx/s, seconds, description, psnr, outputs)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)          # torchrun exports 1; the C search reads it at first use
    from oracle import vnlb_oracle as orc
    import torch
    torch.set_num_threads(threads)
    orc.set_num_threads(threads)
    s = CPU_SAMPLE
    clean, flows = orc.synth_video(s["T"], WORK["H"], WORK["W"], 123, return_flows=True, crop=(s["H"], s["W"]))
    noisy = orc.add_noise(clean, SIGMA, 123)
    torch.manual_seed(123)
    t0 = time.time()
    deno, basic, _ = orc.denoise(noisy, SIGMA, flows=flows)
    dt = time.time() - t0
    px = s["T"] * s["H"] * s["W"]
    desc = ("top-left %dx%dx%d crop of the workload (sigma=10, with its flows), full vnlb.denoise (steps 1+2), "
            "default_params" % (s["W"], s["H"], s["T"]))
    psnr = dict(basic=float(orc.compute_psnrs(basic, clean).mean()), deno=float(orc.compute_psnrs(deno, clean).mean()))
    return px / 1e6 / dt, dt, desc, psnr, dict(clean=clean, noisy=noisy, flows=flows, deno=deno, basic=basic)


def cpu_config1(threads):
    """BASELINE configs[0]: davis_64x64-shaped synthetic RGB, 3 frames, sigma=20, vnlb.denoise on the CPU."""
    from oracle import vnlb_oracle as orc
    import torch
    orc.set_num_threads(threads)
    clean = orc.synth_video(3, 64, 64, 123)
    noisy = orc.add_noise(clean, 20., 123)
    torch.manual_seed(123)
    t0 = time.time()
    deno, basic, _ = orc.denoise(noisy, 20.)
    dt = time.time() - t0
    return dict(workload="64x64x3 synthetic RGB, sigma=20, no flow (BASELINE configs[0])", seconds=dt,
                Mpx_per_s=3 * 64 * 64 / 1e6 / dt, cores=threads, kind="port",
                psnr=dict(noisy=float(orc.compute_psnrs(noisy, clean).mean()), basic=float(orc.compute_psnrs(basic, clean).mean()),
                          deno=float(orc.compute_psnrs(deno, clean).mean()))), (clean, noisy, deno, basic)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, desc, psnr, _ = cpu_reference_run(threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    line = dict(metric=METRIC, value=v, unit="Mpx/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD_DESC + "; each step = bounded sample: " + desc),
                cpu_baseline=dict(value=v, unit="Mpx/s", cores=threads, kind="port", sample=desc),
                e2e=dict(value=v, unit="Mpx/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), psnr=psnr)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vnlb_b200
    from vnlb_b200 import _lib, dist as vdist, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    assert world == n_gpus or world == 1 and n_gpus == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    T, H, W = (args.frames or WORK["T"]), WORK["H"], WORK["W"]
    mpx = T * H * W / 1e6
    clean, flows_np = synth.synth_video(T, H, W, 123, return_flows=True)
    noisy = synth.add_noise(clean, SIGMA, 123)
    noisy_alt = synth.add_noise(clean, SIGMA_ALT, 123)
    if rank != 0:
        clean = None
    pin = lambda a: torch.from_numpy(a).pin_memory()
    noisy_pin = pin(noisy)
    flows_pin = dict(fflow=pin(flows_np["fflow"]), bflow=pin(flows_np["bflow"]))
    noisy_dev = noisy_pin.to(device)
    flows_dev = dict(fflow=flows_pin["fflow"].to(device), bflow=flows_pin["bflow"].to(device))
    if world == 1:
        out_pin = torch.empty_like(noisy_pin).pin_memory()
    else:       # every rank reads back its own band of the estimate (the output stays sharded like the work)
        out_pin = torch.empty((T * 3 * (H // world + H // (4 * world) + 16) * W,), dtype=torch.float32).pin_memory()
    del flows_np

    def call(x, fl, sigma=SIGMA, stats=None, params=None, gather=True):
        params = params if params is not None else vnlb_b200.get_params(sigma)
        if world > 1:
            return vdist.denoise_distributed(x, sigma, flows=fl, params=params, stats=stats, device=device, max_flow=MAX_FLOW,
                                             gather=gather)
        return vnlb_b200.denoise(x, sigma, gpuid=local_rank, verbose=False, flows=fl, schedule="fast", params=params, stats=stats)

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
        finally:
            gc.enable()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    result = {}

    def step_dev():
        result["deno"], result["basic"], _ = call(noisy_dev, flows_dev)

    # ---- warm-up: the same step as the timed one (kernels loaded, caching allocator in its steady state) ----
    for _ in range(args.warmup):
        step_dev()

    # ---- device-resident throughput (`value`) ----
    l0 = int(_lib.lib.vnlb_kernel_launches())   # kernels launched by libvnlb_b200.so, counted in the library
    if rank == 0:   # sample the clocks every 64th C-ABI call of the timed steps: the GPU is busy at those moments
        _lib.on_launch = lambda n: sampler.sample() if n % 64 == 0 else None
    ms = timed(step_dev, args.steps, "value")
    launches = (int(_lib.lib.vnlb_kernel_launches()) - l0) // max(args.steps, 1)

    # ---- end to end through the public API with HOST buffers (`e2e`): video + flows host->device (each rank only its
    #      band + halo rows), result device->host on rank 0, all inside the timed region ----
    st_e2e = {}

    def step_e2e():
        if world == 1:
            d, _, _ = call(noisy_pin, flows_pin, stats=st_e2e)
            out_pin.copy_(d, non_blocking=True)
        else:
            d, _, _ = call(noisy_pin, flows_pin, stats=st_e2e, gather=False)
            out_pin[:d.numel()].copy_(d.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, "e2e")
    _lib.on_launch = None
    clocks = sampler.summary() if rank == 0 else None
    rows_copied = st_e2e.get("layout", {}).get("rows_copied", H)
    h2d = torch.tensor([float(rows_copied) * W * T * (3 + 4) * 4], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(h2d)
    h2d_bytes = int(h2d.item())

    # ---- the other noise level of configs[4] ----
    noisy_alt_dev = torch.from_numpy(noisy_alt).to(device)
    call(noisy_alt_dev, flows_dev, SIGMA_ALT)
    st_alt = {}
    n_alt = max(1, min(args.steps, 3))
    ms_alt = timed(lambda: result.__setitem__("alt", call(noisy_alt_dev, flows_dev, SIGMA_ALT, st_alt)), n_alt, "sigma50")
    deno_alt = result.pop("alt")[0]

    # ---- per-stage device time (CUDA events on the launching streams) -> roofline of the dominant stage ----
    _lib.timer = _lib.StageTimer()
    st2 = {}
    call(noisy_dev, flows_dev, stats=st2)
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
                        exchange_ms_rank0=[round(x, 2) for x in st2.get("exchange_ms", [])],
                        exchange_bytes_rank0=st2.get("exchange_bytes"), layout_rank0=st2.get("layout"),
                        rebalance_rank0=st2.get("rebalance"))

    if rank == 0:
        hbm_peak, hbm_kind, fp32_peak, fp32_kind, fp32_rec = load_peaks()
        psnr_of = lambda x, ref=clean: float(vnlb_b200.compute_psnrs(x.cpu().numpy() if torch.is_tensor(x) else x, ref).mean())
        psnr = dict(noisy=psnr_of(noisy), basic=psnr_of(result["basic"]), deno=psnr_of(result["deno"]))
        psnr_alt = dict(noisy=psnr_of(noisy_alt), deno=psnr_of(deno_alt))
        psnr_delta = {}
        if world > 1:
            # the same video on ONE GPU (outside the timed region): multi-GPU output within the PSNR tolerance of it
            d1, b1, _ = vnlb_b200.denoise(noisy_dev, SIGMA, gpuid=local_rank, verbose=False, flows=flows_dev, schedule="fast")
            psnr_delta["vs_n1_deno_db"] = psnr["deno"] - psnr_of(d1)
            psnr_delta["vs_n1_basic_db"] = psnr["basic"] - psnr_of(b1)
            del d1, b1
        # stage keys are "<stage>_s<step>"
        merged = {}
        for k, v in stage.items():
            base = k.rsplit("_s", 1)[0]
            m = merged.setdefault(base, dict(ms=0.0, launches=0))
            m["ms"] += v["ms"]
            m["launches"] += v["launches"]
        step_ms = sum(v["ms"] for v in merged.values())
        # Algorithmic work per group (SURVEY 8d / DESIGN.md section 5), per VNLB step: k = 100 / 60, p = 98, C = 3.
        #   fused Bayes: nominal LAPACK-style flop count 35.8 / 31.7 MFLOP; bytes = gather k*D*4 (noisy; + basic in step 2)
        #   + scatter k*(D+p)*4.  search: N_cand*p*C_d*3 flops.
        bayes_flops, bayes_bytes = [35.8e6, 31.7e6], [100 * 294 * 4 + 100 * 392 * 4, 2 * 60 * 294 * 4 + 60 * 392 * 4]
        traffic_rec = load_ncu_traffic()
        roof = None
        dom = "bayes_s0" if "bayes_s0" in stage else None
        if dom:
            d = stage[dom]
            sec = d["ms"] / 1e3
            flops = bayes_flops[0] * ngroups[0]
            byts = bayes_bytes[0] * ngroups[0]
            per_launch_groups = ngroups[0] / max(d["launches"], 1)
            traffic = None
            if traffic_rec and "bayes_step1_dram_bytes_per_group" in traffic_rec:
                traffic = traffic_rec["bayes_step1_dram_bytes_per_group"] * per_launch_groups
            roof = dict(
                kernel="vnlb_bayes_aggregate_fused, VNLB step 1 (cov_tridiag_kernel + tridiag_tail2_kernel + tridiag_tail_kernel + bayes_kernel<fused,split>): "
                       "gather + covariance + eigen-decomposition + Wiener filter + aggregation of one round of groups",
                bound="fp32", achieved=flops / sec / 1e12, peak=fp32_peak, unit="TFLOP/s", frac=flops / sec / 1e12 / fp32_peak,
                peak_kind=fp32_kind, traffic=traffic,
                traffic_source=(traffic_rec or {}).get("source", None),
                algorithmic_flops_per_launch=bayes_flops[0] * per_launch_groups,
                algorithmic_bytes_per_launch=bayes_bytes[0] * per_launch_groups,
                launches=d["launches"], avg_launch_ms=d["ms"] / max(d["launches"], 1),
                launch_unit="one C-ABI call = one round of groups (4 kernels)", share_of_step=d["ms"] / max(step_ms, 1e-9),
                flop_count="nominal LAPACK-style count of SURVEY 8d (cov 2kp^2C + eig 9p^3C + filter 4kprC = 35.8 MFLOP per group); "
                           "the kernels execute fewer: only the eigenpairs above the Wiener threshold are computed",
                ncu=(traffic_rec or {}).get("bayes_step1_pipes", None),
                hbm=dict(achieved=byts / sec / 1e9, peak=hbm_peak, unit="GB/s", frac=byts / sec / 1e9 / hbm_peak, peak_kind=hbm_kind,
                         note="not the binding roof: arithmetic intensity ~130 flop/B against a ridge of ~10"),
                step2=dict(achieved=bayes_flops[1] * ngroups[1] / (stage["bayes_s1"]["ms"] / 1e3) / 1e12,
                           frac=bayes_flops[1] * ngroups[1] / (stage["bayes_s1"]["ms"] / 1e3) / 1e12 / fp32_peak,
                           note="nominal count of the p x p problem; the kernel solves the 60 x 60 Gram problem (SURVEY: ~12 MFLOP)")
                if "bayes_s1" in stage else None)
        srch = {}
        for s_ in (0, 1):
            key = "search_s%d" % s_
            if key in stage and ngroups[s_]:
                dc = 1 if s_ == 0 else 3
                q = ngroups[s_] + (st2.get("ndropped", [0, 0])[s_] if st2.get("ndropped") else 0)
                tf = 9477 * 98 * dc * 3 * q / (stage[key]["ms"] / 1e3) / 1e12
                srch["step%d" % (s_ + 1)] = dict(queries=q, ms=stage[key]["ms"], Mqueries_per_s=q / stage[key]["ms"] / 1e3,
                                                 algorithmic_tflops=tf, frac_of_fp32_peak=tf / fp32_peak)
        extras = {}
        cpu = None
        if n_gpus == 1 and not args.quick:
            extras, cpu = single_gpu_extras(vnlb_b200, device, fp32_peak, hbm_peak, psnr_delta, args)
        line = dict(
            metric=METRIC, value=mpx * args.steps / (ms / 1e3), unit="Mpx/s", n_gpus=n_gpus, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong",
            vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=WORKLOAD_DESC if T == WORK["T"] else WORKLOAD_DESC + " [frames overridden: %d]" % T,
                        params="default_params: 7x7x2 patches, 27x27 window, +-6 frames, k=100/60, rank 39",
                        schedule="fast (device-side rounds, in-round conflict resolution)",
                        parallelism="1 GPU" if n_gpus == 1 else "%d row bands + halo, border-strip exchange (NCCL send/recv)" % n_gpus,
                        l2="working set (noisy + flows + basic + accumulators ~%.1f GB) exceeds the 126 MB L2" % (T * H * W * 4 * 14 / 1e9)),
            e2e=dict(value=mpx * args.steps / (ms_e2e / 1e3), unit="Mpx/s", ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=int(noisy.nbytes),
                     note="video + forward/backward flows host->device (each rank its band + halo rows only); estimate device->host, at "
                          "N > 1 every rank its own band (the output stays sharded like the work; `value` includes the all-gather "
                          "that gives every rank the full frames)"),
            sigma50=dict(value=mpx * n_alt / (ms_alt / 1e3), unit="Mpx/s", ms_per_step=ms_alt / n_alt, steps=n_alt,
                         groups_per_step=[int(g) for g in st_alt.get("ngroups", [])][-2:], psnr=psnr_alt),
            gpu_launches=int(launches), clocks=clocks, roofline=roof, search=srch, cpu_baseline=cpu,
            stages_ms={k: round(v["ms"], 3) for k, v in sorted(stage.items())}, groups_per_step=ngroups,
            dropped_per_step=st2.get("ndropped"), psnr=psnr, psnr_delta=psnr_delta or None,
            rounds=st2.get("nrounds"), per_rank=per_rank, per_step_ms_rank0=per_step, **extras)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def single_gpu_extras(vnlb_b200, device, fp32_peak, hbm_peak, psnr_delta, args):
    """N = 1 only, outside the timed regions: the remaining BASELINE configurations and the parity evidence of the benched
    schedule.  Returns (extra keys of the JSON line, cpu_baseline)."""
    import torch
    from vnlb_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import microbench
    dev = str(device)
    psnr_of = lambda x, ref: float(vnlb_b200.compute_psnrs(x.cpu().numpy() if torch.is_tensor(x) else x, ref).mean())
    extras = {}
    torch.cuda.empty_cache()
    # ---- BASELINE configs[1]: 854x480x20, sigma=20, no flow ----
    c2 = CFG2
    clean2 = synth.synth_video(c2["T"], c2["H"], c2["W"], 123)
    noisy2 = torch.from_numpy(synth.add_noise(clean2, c2["sigma"], 123)).to(device)
    for _ in range(3):      # warm-up WITH the results held like in the timed loop: a call whose predecessor's outputs are
        d2, b2, _ = vnlb_b200.denoise(noisy2, c2["sigma"], gpuid=device.index, verbose=False)   # still alive needs fresh blocks (one-time cudaMalloc, +150 ms)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n2 = 3
    import gc
    gc.collect()
    gc.disable()            # as in timed(): a full collection inside a 220 ms call is a host stall, not GPU time
    try:
        torch.cuda.synchronize()
        evs[0].record()
        per_call = [torch.cuda.Event(enable_timing=True) for _ in range(n2)]
        for i2 in range(n2):
            st = {}
            d2, b2, _ = vnlb_b200.denoise(noisy2, c2["sigma"], gpuid=device.index, verbose=False, stats=st)
            per_call[i2].record()
        evs[1].record()
        torch.cuda.synchronize()
    finally:
        gc.enable()
    per_call_ms = [round((evs[0] if i2 == 0 else per_call[i2 - 1]).elapsed_time(per_call[i2]), 1) for i2 in range(n2)]
    ms2 = evs[0].elapsed_time(evs[1]) / n2
    extras["config2"] = dict(workload="854x480x20 synthetic RGB, sigma=20, no flow (BASELINE configs[1])",
                             value=c2["T"] * c2["H"] * c2["W"] / 1e6 / (ms2 / 1e3), unit="Mpx/s", ms_per_step=ms2, steps=n2,
                             per_call_ms=per_call_ms,
                             groups_per_step=st["ngroups"], dropped_per_step=st.get("ndropped"),
                             psnr=dict(basic=psnr_of(b2, clean2), deno=psnr_of(d2, clean2)))
    del noisy2, d2, b2
    # ---- benched schedule vs the reference-exact schedule on a sub-video of the workload (sigma=10, with flows) ----
    s = PARITY_SUB
    cs, fs = synth.synth_video(s["T"], WORK["H"], WORK["W"], 123, return_flows=True, crop=(s["H"], s["W"]))
    ns = synth.add_noise(cs, SIGMA, 123)
    torch.manual_seed(123)
    sp, sf = {}, {}
    dp, bp, _ = vnlb_b200.denoise(ns, SIGMA, gpuid=device.index, verbose=False, flows=fs, schedule="parity", stats=sp)
    df, bf, _ = vnlb_b200.denoise(ns, SIGMA, gpuid=device.index, verbose=False, flows=fs, schedule="fast", stats=sf)
    psnr_delta["fast_vs_parity_schedule"] = dict(
        video="top-left %dx%dx%d crop of the workload" % (s["W"], s["H"], s["T"]),
        deno_db=psnr_of(df, cs) - psnr_of(dp, cs), basic_db=psnr_of(bf, cs) - psnr_of(bp, cs),
        groups_parity=sp["ngroups"], groups_fast=sf["ngroups"], dropped_fast=sf.get("ndropped"))
    del dp, bp, df, bf
    # ---- CPU baseline (oracle port, all host cores) on a bounded sample + the benched schedule on the same sample ----
    threads = host_threads()
    cpu = None
    if not args.no_cpu_baseline:
        v, dt, desc, opsnr, o = cpu_reference_run(threads)
        cpu = dict(value=v, unit="Mpx/s", cores=threads, kind="port", sample=desc, seconds=dt, psnr=opsnr)
        dg, bg, _ = vnlb_b200.denoise(o["noisy"], SIGMA, gpuid=device.index, verbose=False, flows=o["flows"], schedule="fast")
        torch.manual_seed(123)
        dq, bq, _ = vnlb_b200.denoise(o["noisy"], SIGMA, gpuid=device.index, verbose=False, flows=o["flows"], schedule="parity")
        psnr_delta["vs_cpu_oracle_on_cpu_sample"] = dict(
            fast_deno_db=psnr_of(dg, o["clean"]) - opsnr["deno"], fast_basic_db=psnr_of(bg, o["clean"]) - opsnr["basic"],
            parity_deno_db=psnr_of(dq, o["clean"]) - opsnr["deno"], parity_basic_db=psnr_of(bq, o["clean"]) - opsnr["basic"],
            parity_max_abs_deno=float(np.abs(dq.cpu().numpy() - o["deno"]).max()),
            parity_max_abs_basic=float(np.abs(bq.cpu().numpy() - o["basic"]).max()))
        # ---- BASELINE configs[0]: 64x64x3 on the CPU, and the CUDA path on the same input ----
        c1, (cl1, no1, od1, ob1) = cpu_config1(threads)
        torch.manual_seed(123)
        d1, b1, _ = vnlb_b200.denoise(no1, 20., gpuid=device.index, verbose=False, schedule="parity")
        c1["gpu_parity_schedule_max_abs"] = dict(deno=float(np.abs(d1.cpu().numpy() - od1).max()),
                                                 basic=float(np.abs(b1.cpu().numpy() - ob1).max()))
        extras["config1_cpu"] = c1
    # ---- BASELINE configs[2] / configs[3]: microbenchmarks on the shipped kernels ----
    torch.cuda.empty_cache()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    extras["config3_search"] = microbench.search_microbench(dev, 65536, fp32_peak, flush)
    extras["config4_bayes"] = microbench.bayes_microbench(dev, 16384, fp32_peak, hbm_peak, flush)
    return extras, cpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the N = 1 extras (other BASELINE configs, parity evidence, CPU baseline)")
    ap.add_argument("--frames", type=int, default=None, help="override frame count (debugging only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
