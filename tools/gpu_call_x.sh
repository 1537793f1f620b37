#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c29
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python tools/bench_loss.py 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3),"roofline",d["roofline"]["avg_launch_us"],d["roofline"]["frac"])
PY
