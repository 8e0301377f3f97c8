"""Runs tools/fp32_peak.cu on the GPU and writes profiles/<out>.json (default profiles/fp32_peak.json, tracked): the
measured FP32 CUDA-core peak bench.py divides by.  The binary is cross-compiled in the build container
(`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp32_peak tools/fp32_peak.cu`, build/ travels with gpurun);
it is compiled here if missing.   usage: python tools/measure_fp32_peak.py [out.json]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "build", "fp32_peak")
if not os.path.exists(exe):
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe,
                           os.path.join(ROOT, "tools", "fp32_peak.cu")])
rec = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip().splitlines()[0]
    rec["nvidia_smi_after"] = q
except Exception:
    pass
rec["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "fp32_peak.json")
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec))
