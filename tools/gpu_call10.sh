#!/bin/bash
cd "$(dirname "$0")/.."
for ns in 0 4000 9000 14000 20000; do
  echo "== stagger $ns"
  VNLB_SEARCH_STAGGER_NS=$ns timeout 300 python tools/search_ab.py 16384 2> gpurun_out/r2_search_ab.err | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items(): print(k, {p:(x['ms'],x.get('same_as_path1')) for p,x in v.items() if p!='path2'})
"
done
