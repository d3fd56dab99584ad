#!/bin/bash
# r02b: smoke (N=2^16 + m=4096), gpu tests (incl. world-2-on-one-GPU sharded parity), full bench line with the key-switch blocks,
# reference arm, ncu launch list + full capture of one step.
TAG=${1:-r02b}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | cut -c1-1200
tail -5 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee $OUT/bench_ref_$TAG.json | cut -c1-400
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
   python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu --no-e2e --no-ks > $OUT/ncu_bench_$TAG.log 2>&1
tail -2 $OUT/ncu_bench_$TAG.log | cut -c1-300
echo "== ncu full: one whole step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_conv|k1_fwd_blk|k1_inv_blk|k1_ks_inner|k_pw' -s 51 -c 17 -f -o $OUT/prof_$TAG \
   python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu --no-e2e --no-ks > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-300
ls -la $OUT
fi
