#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c21
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; python - <<PY
import json
d=json.loads(open("${O}_bench.txt").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"])
PY
tail -3 ${O}_bench.txt | head -2 | cut -c1-300
FSNET_WGRAD_STREAM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_nostream.txt 2>&1; tail -1 ${O}_bench_nostream.txt | cut -c1-240
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > ${O}_tests_all.txt; tail -3 ${O}_tests_all.txt
