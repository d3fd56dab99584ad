#!/bin/bash
# lean N-GPU run: bench.py under torchrun with the multiply block, a small config-3 block and the sharded block.  usage: gpu_multi_lean.sh N [TAG]
N=${1:-4}; TAG=${2:-r02m$N}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --ks-count 256 --ks-steps 2 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $OUT/bench_$TAG.err | tail -8 | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_$TAG.json") if l.startswith("{")][-1])
    print("mult", round(d["value"],1), "ks", round(d["keyswitch"]["value"],1))
    s = d["sharded_keyswitch"]
    print("sharded pure best", round(s["value"],1), "bit_exact", s["bit_exact_vs_unsharded"], "best grouping", s["best_grouping"])
    for m, r in s["modes"].items():
        print("  ", m, round(r["value"],1), "ms/step", round(r["ms_per_step"],3), "graph", r["cuda_graph"], "launches", r["gpu_launches_per_step"], "engine ms", r["engine_kernel_ms_per_step"], "exact", r["bit_exact_vs_unsharded"])
        for k in r["phase_ms"][:9]: print("      ", k)
except Exception as e:
    print("parse failed", e)
PY
