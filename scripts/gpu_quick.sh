#!/bin/bash
# quick A/B on the GPU: parity subset + bench kernel table.  usage: gpu_quick.sh TAG [ENV=VAL ...]
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_engine_parity.py tests/test_engine_semantics.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_$TAG.log
for cfg in "$@" "X=0"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-ks 2>$OUT/err_$TAG.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
for k in d['kernels']: print('  ', k)
" | tee -a $OUT/sweep_$TAG.txt
  tail -3 $OUT/err_$TAG.txt
done
