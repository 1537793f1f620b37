#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c9
timeout 1500 python -m pytest tests -q -m gpu -rA -s --timeout 900 -p no:cacheprovider > ${O}_gpu_tests.txt 2>&1
echo "gpu rc=$?" >> ${O}_gpu_tests.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench_default.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --e2e-sync 1 --no-cpu-baseline > ${O}_bench_e2esync.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file ${O}_launches.csv python tools/profile_step.py > ${O}_ncu.log 2>&1
grep -E "passed|failed" ${O}_gpu_tests.txt | tail -3
grep -E "FAILED|ERROR|gradient tensors" ${O}_gpu_tests.txt | cut -c1-250 | head -20
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_*.txt
grep -o '"e2e": {[^}]*}' ${O}_bench_*.txt
