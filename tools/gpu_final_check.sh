#!/bin/bash
# final N = 1 check: full GPU test-suite, smoke, default bench
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
SECONDS=0
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_e.json") if l.startswith("{")][-1])
print("N=1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["sigma50"]["value"], d["config2"]["value"], d["config2"]["per_call_ms"], d["roofline"]["frac"], d["stages_ms"], d["config4_bayes"]["step1"]["ms"], d["config4_bayes"]["step2"]["ms"], d["psnr_delta"]["fast_vs_parity_schedule"])
PY
