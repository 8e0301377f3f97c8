#!/bin/bash
# default bench at N = 1 (new search kernel)
cd "$(dirname "$0")/.."
SECONDS=0
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_b.json") if l.startswith("{")][-1])
    print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages_ms"], d["sigma50"]["value"], d["config2"]["value"], d["config3_search"], d["psnr_delta"])
except Exception as e:
    print("bench unreadable", e)
PY
