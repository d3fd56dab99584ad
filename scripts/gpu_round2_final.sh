#!/bin/bash
# Final 1-GPU session of round 2: smoke, all gpu tests, the driver-shaped bench line, the reference arm, ncu launch list and a full
# ncu capture of one step.   usage (through gpurun): bash scripts/gpu_round2_final.sh [tag]
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | cut -c1-700
tail -5 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tee $OUT/bench_ref_$TAG.json | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ks > $OUT/ncu_bench_$TAG.log 2>&1
tail -1 $OUT/ncu_bench_$TAG.log | cut -c1-200
echo "== ncu full: one whole step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_conv|k2_fwd_blk|k1_inv_blk|k1_ks_inner|k_pw' -s 51 -c 17 -f -o $OUT/prof_$TAG \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-ks > $OUT/ncu_full_$TAG.log 2>&1
tail -1 $OUT/ncu_full_$TAG.log | cut -c1-200
ls -la $OUT | tail -12
