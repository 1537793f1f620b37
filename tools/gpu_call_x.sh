#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c23
for v in "FSNET_WGRAD_STREAMS=1" "FSNET_WGRAD_STREAMS=2" "FSNET_WGRAD_STREAMS=3"; do
env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
print("$v value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3))
PY
done
