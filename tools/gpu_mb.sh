#!/bin/bash
cd "$(dirname "$0")/.."
python tools/microbench.py 4096 16384 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print({k: round(v['ms'],3) for k,v in d['bayes_config4'].items()})"
