#!/bin/bash
# rebalancing experiment at N GPUs: args N round damp
N=$1
cd "$(dirname "$0")/.."
export VNLB_REBALANCE=1 VNLB_REBALANCE_ROUND=$2 VNLB_REBALANCE_DAMP=$3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_reb.json 2> gpurun_out/r2_bench_reb.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_reb.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["psnr_delta"], d["per_rank"]["groups"], d["per_rank"]["stage_ms"], d["per_rank"]["value_step_ms"][0], d["per_rank"]["exchange_ms_rank0"], d["per_rank"].get("rebalance_rank0"))
PY
