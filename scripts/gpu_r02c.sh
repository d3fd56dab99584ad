#!/bin/bash
# r02c: first run of the TMA-staged blk kernels (k2_*) on the GPU: parity, then A/B against the cp.async kernels.
OUT=gpurun_out; mkdir -p $OUT
echo "== parity (k2 kernels)"; timeout 900 python -m pytest tests/test_engine_parity.py tests/test_engine_semantics.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_r02c.log
for v2 in 1 0; do
  echo "== bench HB_BLK_V2=$v2"
  HB_BLK_V2=$v2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-ks 2>$OUT/err_r02c.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
for k in d['kernels']: print('  ', k)
" | tee -a $OUT/sweep_r02c.txt
  tail -3 $OUT/err_r02c.txt
done
