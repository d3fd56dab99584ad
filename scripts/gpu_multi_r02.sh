#!/bin/bash
# N-GPU run of the driver's command line: bench.py under torchrun (replicas for the multiply and config-3 blocks, prime-sharded
# key switch for config 4), plus the sharded NCCL tests.   usage: gpu_multi_r02.sh N [TAG]
N=${1:-2}; TAG=${2:-r02m$N}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$TAG.txt 2>&1
echo "== sharded tests"; timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_sharded_$TAG.log
echo "== bench --gpus $N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "rc=$?"; tail -5 $OUT/bench_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_$TAG.json") if l.startswith("{")][-1])
    print("mult", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ks", round(d["keyswitch"]["value"],1))
    s = d["sharded_keyswitch"]
    print("sharded best", round(s["value"],1), "bit_exact", s["bit_exact_vs_unsharded"])
    for m, r in s["modes"].items():
        print("  ", m, round(r["value"],1), "ms/step", round(r["ms_per_step"],3), "graph", r["cuda_graph"], "launches", r["gpu_launches_per_step"], "engine ms", r["engine_kernel_ms_per_step"])
        for k in r["phase_ms"][:12]: print("      ", k)
except Exception as e:
    print("parse failed", e)
PY
