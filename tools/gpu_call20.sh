#!/bin/bash
# 2 GPUs: NCCL tests of the band + halo path, then the bench at N = 2
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/r2_pytest_dist.log
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; echo "bench2 rc=$? wall=${SECONDS}s"; tail -2 gpurun_out/r2_bench_n2_c.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_c.json") if l.startswith("{")][-1])
    print("N=2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["sigma50"]["value"], d["psnr_delta"], d["per_rank"])
except Exception as e:
    print("bench2 unreadable", e)
PY
