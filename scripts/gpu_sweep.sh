#!/bin/bash
./bench_micro/ubench_butterfly | grep -v mac
for cfg in "16 8 0" "8 8 0" "16 8 1"; do
  set -- $cfg
  echo "== HB_CHUNK=$1 batch=$2 NO_SPECIAL=$3"
  if [ "$3" = "1" ]; then export HB_NO_SPECIAL=1; else unset HB_NO_SPECIAL; fi
  HB_CHUNK=$1 timeout 200 python bench.py --steps 20 --warmup 3 --batch $2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'mult/s', 'ms/step',round(d['ms_per_step'],3), 'clocks',d['clocks'].get('sm_mhz'), [(k['kernel'],k['ms']) for k in d['kernels'][:5]])"
done
