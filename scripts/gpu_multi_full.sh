#!/bin/bash
# the driver's N-GPU command with default blocks (e2e, 1024-ciphertext config 3, sharded modes): usage gpu_multi_full.sh N [TAG]
N=${1:-8}; TAG=${2:-r02f$N}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $OUT/bench_$TAG.err | tail -8 | cut -c1-300
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_$TAG.json") if l.startswith("{")][-1])
print("mult", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"], "ks", round(d["keyswitch"]["value"],1))
s = d["sharded_keyswitch"]
print("sharded pure best", round(s["value"],1), "bit_exact", s["bit_exact_vs_unsharded"], s["best_grouping"])
for m, r in s["modes"].items():
    print("  ", m, round(r["value"],1), "ms/step", round(r["ms_per_step"],3), "exact", r["bit_exact_vs_unsharded"])
print("general_m", d.get("general_m"))
print("clocks", d["clocks"])
PY
