#!/bin/bash
# Multi-GPU session (run with gpurun --gpus N): sharded key-switch parity over NCCL, scaling of both modes.
TAG=${1:-r01m}; N=${2:-2}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus_$TAG.txt
echo "== pytest sharded nccl"; timeout 300 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_sharded_$TAG.log
for n in 1 2 $N; do
  if [ $n -gt $N ]; then continue; fi
  if [ $n -eq 2 ] && [ $N -eq 2 ] && [ "$seen2" = "1" ]; then continue; fi
  if [ $n -eq 2 ]; then seen2=1; fi
  for mode in sharded replicas; do
    echo "== keyswitch $mode gpus=$n"
    if [ $n -eq 1 ]; then timeout 150 python bench_keyswitch.py --mode $mode --steps 10 --batch ${BATCH:-8} 2>$OUT/ks_${mode}_${n}_$TAG.err | tee $OUT/ks_${mode}_${n}_$TAG.json | cut -c1-400
    else timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench_keyswitch.py --gpus $n --mode $mode --steps 10 --batch ${BATCH:-8} 2>$OUT/ks_${mode}_${n}_$TAG.err | tee $OUT/ks_${mode}_${n}_$TAG.json | cut -c1-400; fi
    tail -3 $OUT/ks_${mode}_${n}_$TAG.err
  done
done
echo "== bench.py gpus=$N"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 2>$OUT/bench_${N}_$TAG.err | tee $OUT/bench_${N}_$TAG.json | cut -c1-600
tail -3 $OUT/bench_${N}_$TAG.err
