#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c27
for v in "FSNET_X=1" "FSNET_SMOOTH_STREAM=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
try:
    d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
    print("$v value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print("$v failed"); print(open("${O}_bench.txt").read()[-2000:])
PY
done
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -3
