#!/bin/bash
cd "$(dirname "$0")/.."
SECONDS=0
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_d.json") if l.startswith("{")][-1])
print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config2"])
PY
