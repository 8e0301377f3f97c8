#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "round_draw or graph_rounds or deterministic" > gpurun_out/r2_pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_graph.log
timeout 600 python tools/graph_ab.py 2>&1 | tail -12
