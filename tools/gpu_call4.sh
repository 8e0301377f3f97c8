#!/bin/bash
# round-2 GPU call: tests, A/B of the two-row elimination, bench at N = 2, ncu captures (1 GPU)
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -12 gpurun_out/r2_pytest3.log
VNLB_TAIL2=0 python tools/microbench.py 4096 16384 > gpurun_out/r2_mb_tail1b.json 2>&1
VNLB_TAIL2=1 python tools/microbench.py 4096 16384 > gpurun_out/r2_mb_tail2b.json 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_mb_tail1b.json", "gpurun_out/r2_mb_tail2b.json"):
    try:
        d = json.load(open(f)); print(f, {k: round(v["ms"], 3) for k, v in d["bayes_config4"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2_bench_n2_b.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_b.json") if l.startswith("{")][-1])
    print("N=2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["psnr_delta"], d["per_rank"])
except Exception as e:
    print("bench2 unreadable", e)
PY
# ncu: per-kernel counters of one round of 2048 groups per step, then the launch list of a bench call
python tools/run_kernels.py fused 2048 > gpurun_out/r2_plain_kernels.log 2>&1 && \
ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r2a -f python tools/run_kernels.py fused 2048 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py --quick --frames 8 --steps 1 --warmup 1 > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --quick --frames 8 --steps 1 --warmup 1 > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
