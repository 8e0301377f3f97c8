#!/bin/bash
# final N = 1 check: full GPU test-suite, smoke, default bench
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
SECONDS=0
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$? wall=${SECONDS}s"
SECONDS=0
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$? wall=${SECONDS}s"; tail -1 gpurun_out/r2_bench_ref.json | cut -c1-600
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n1_c.json") if l.startswith("{")][-1])
    print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages_ms"], d["sigma50"]["value"], d["config2"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["search"], d["config4_bayes"])
except Exception as e:
    print("bench unreadable", e)
PY
