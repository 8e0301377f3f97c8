#!/bin/bash
cd "$(dirname "$0")/.."
python tools/dbg_search.py 2 > gpurun_out/dbg2.log 2>&1; echo "path2 rc=$?"; tail -2 gpurun_out/dbg2.log
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_search.py 0 > gpurun_out/dbg0.log 2>&1; echo "path0 rc=$?"; grep -v "^$" gpurun_out/dbg0.log | head -60
