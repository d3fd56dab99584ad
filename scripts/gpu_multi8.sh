#!/bin/bash
TAG=$1; N=${2:-8}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== parity p2p $N ranks"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29720 tests/mp/sharded_worker.py cuda 131072,257,1,400,2 p2p 2>&1 | grep "RANK" | sort | tr '\n' ' '; echo
for ex in p2p gather; do
  echo "== sharded cfg4 batch=16 gpus=$N exchange=$ex"
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench_keyswitch.py --gpus $N --mode sharded --exchange $ex --steps 10 --batch 16 --profile 2>$OUT/ks_sh_${ex}_${N}_$TAG.err | grep '^{' | tee $OUT/ks_sh_${ex}_${N}_$TAG.json | cut -c1-200
done
echo "== replicas cfg3 gpus=$N"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench_keyswitch.py --gpus $N --mode replicas --steps 10 --batch 8 2>$OUT/ks_rep_${N}_$TAG.err | grep '^{' | tee $OUT/ks_rep_${N}_$TAG.json | cut -c1-200
echo "== bench.py gpus=$N"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/bench_${N}_$TAG.err | grep '^{' | tee $OUT/bench_${N}_$TAG.json | cut -c1-300
