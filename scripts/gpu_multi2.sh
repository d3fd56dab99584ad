#!/bin/bash
# sharded key-switch scaling at a larger batch: bash scripts/gpu_multi2.sh TAG N BATCH
TAG=$1; N=$2; B=${3:-32}; OUT=gpurun_out; mkdir -p $OUT
echo skip-pytest
for ex in p2p gather; do
for n in 1 $N; do
  if [ $n -eq 1 ] && [ $ex = gather ]; then continue; fi
  echo "== sharded batch=$B gpus=$n exchange=$ex"
  if [ $n -eq 1 ]; then timeout 150 python bench_keyswitch.py --mode sharded --exchange $ex --steps 10 --batch $B --profile 2>$OUT/ks_sh_${ex}_b${B}_${n}_$TAG.err | tee $OUT/ks_sh_${ex}_b${B}_${n}_$TAG.json | cut -c1-200
  else timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29710 bench_keyswitch.py --gpus $n --mode sharded --exchange $ex --steps 10 --batch $B --profile 2>$OUT/ks_sh_${ex}_b${B}_${n}_$TAG.err | grep '^{' | tee $OUT/ks_sh_${ex}_b${B}_${n}_$TAG.json | cut -c1-200; fi
done
done
