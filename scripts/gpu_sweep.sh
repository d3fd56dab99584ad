#!/bin/bash
OUT=gpurun_out; TAG=${1:-sweep}
for cfg in "16 8" "8 8" "16 16" "8 16" "4 8"; do
  set -- $cfg
  echo "== HB_CHUNK=$1 batch=$2"
  HB_CHUNK=$1 timeout 200 python bench.py --steps 20 --warmup 3 --batch $2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'mult/s', 'ms/step',round(d['ms_per_step'],3), 'clocks',d['clocks'].get('sm_mhz'), [(k['kernel'],k['ms']) for k in d['kernels'][:5]])"
done
