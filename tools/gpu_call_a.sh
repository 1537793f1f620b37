#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c8
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench_default.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --e2e-sync 1 --no-cpu-baseline > ${O}_bench_e2esync.txt 2>&1
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd -s 76 -c 12 -o ${O}_bnbwd python tools/profile_step.py > ${O}_ncu.log 2>&1
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_*.txt
grep -o '"e2e": {[^}]*}' ${O}_bench_*.txt
