#!/bin/bash
# r02a: pipe micro-benchmarks, full gpu tests, conv1 on the GPU, CONV1 / CHUNK sweeps of the bench.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_r02a.txt 2>&1
echo "== ubench pipes"; timeout 300 bench_micro/ubench_pipes 2>&1 | tee $OUT/ubench_pipes_r02a.txt | tail -45
echo "== ubench butterfly"; timeout 300 bench_micro/ubench_butterfly 2>&1 | tee $OUT/ubench_butterfly_r02a.txt | tail -12
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_r02a.log
for c1 in 0 1; do for ch in 32 16 8; do
  echo "== bench CONV1=$c1 CHUNK=$ch"
  HB_CONV1=$c1 HB_CHUNK=$ch timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>$OUT/err.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
for k in d['kernels']: print('  ', k)
" | tee -a $OUT/sweep_r02a.txt
done; done
