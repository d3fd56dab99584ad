#!/bin/bash
# general-m check on the GPU: Bluestein parity tests (incl. config 5's ring at full size) + the row-level bench.  usage: gpu_genm.sh TAG
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
echo "== parity"; timeout 600 python -m pytest tests/test_general_m.py tests/test_engine_parity.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_$TAG.log
echo "== bench"; timeout 200 python bench_general_m.py 2>$OUT/err_$TAG.txt | tee $OUT/genm_$TAG.json
tail -3 $OUT/err_$TAG.txt
