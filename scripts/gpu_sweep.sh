#!/bin/bash
for b in 16 32; do
  echo "== batch=$b"
  timeout 200 python bench.py --steps 15 --warmup 3 --batch $b --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'mult/s', 'ms/step',round(d['ms_per_step'],3), [(k['kernel'],k['ms']) for k in d['kernels'][:5]])"
done
