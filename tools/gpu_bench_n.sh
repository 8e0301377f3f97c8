#!/bin/bash
# bench at N GPUs (argument), final code
N=$1
cd "$(dirname "$0")/.."
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n${N}_c.json 2> gpurun_out/r2_bench_n${N}_c.err; echo "bench$N rc=$? wall=${SECONDS}s"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n%s_c.json" % N) if l.startswith("{")][-1])
    print("N=" + N, d["value"], d["ms_per_step"], d["e2e"]["value"], d["sigma50"]["value"], d["psnr_delta"], d["per_rank"]["groups"], d["per_rank"]["stage_ms"], d["per_rank"]["value_step_ms"][0], d["per_rank"]["exchange_ms_rank0"])
except Exception as e:
    print("bench unreadable", e)
PY
