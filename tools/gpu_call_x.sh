#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c30
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_optim_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"weight_planes_batched|wgrad_to_param_batched" -c 6 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "weight_planes_batched|wgrad_to_param_batched|gpu__time_duration" | paste - - | awk '{print $1, $NF}' | head -8
