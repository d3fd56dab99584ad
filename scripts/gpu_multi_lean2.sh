#!/bin/bash
# lean 2-GPU check: sharded tests + the sharded block only (small multiply / config-3 blocks)
N=${1:-2}; TAG=${2:-r02n$N}
OUT=gpurun_out; mkdir -p $OUT
echo "== sharded tests"; timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --ks-count 64 --ks-steps 1 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $OUT/bench_$TAG.err | tail -8 | cut -c1-300
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_$TAG.json") if l.startswith("{")][-1])
s = d["sharded_keyswitch"]
print("sharded pure best", round(s["value"],1), "bit_exact", s["bit_exact_vs_unsharded"])
for m, r in s["modes"].items():
    print("  ", m, round(r["value"],1), "ms/step", round(r["ms_per_step"],3), "launches", r["gpu_launches_per_step"], "engine ms", r["engine_kernel_ms_per_step"], "exact", r["bit_exact_vs_unsharded"])
    for k in r["phase_ms"][:9]: print("      ", k)
PY
