#!/bin/bash
# r02e: batch sweep and ncu full capture with the k2 kernels
OUT=gpurun_out; mkdir -p $OUT
for b in 16 32; do
  echo "== bench batch $b"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-ks --batch $b 2>$OUT/err_r02e.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
for k in d['kernels']: print('  ', k)
" | tee -a $OUT/sweep_r02e.txt
done
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_conv|k2_fwd_blk|k2_inv_blk|k1_ks_inner|k_pw' -s 51 -c 17 -f -o $OUT/prof_r02e \
   python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu --no-e2e --no-ks > $OUT/ncu_full_r02e.log 2>&1
tail -2 $OUT/ncu_full_r02e.log | cut -c1-300
