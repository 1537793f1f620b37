#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c28
for v in "FSNET_X=1" "FSNET_FUSE_BN_ACT=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
try:
    d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
    print("$v value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print("$v failed"); print(open("${O}_bench.txt").read()[-2000:])
PY
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > ${O}_tests_all.txt; tail -3 ${O}_tests_all.txt
