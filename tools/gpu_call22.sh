#!/bin/bash
N=$1
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_probe.py 8 2>&1 | grep "gather=" | tail -24
