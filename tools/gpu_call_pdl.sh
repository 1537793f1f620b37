#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c20
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; echo "$(tail -1 ${O}_bench.txt | cut -c1-260)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_ncu.log 2>&1; tail -2 ${O}_ncu.log | cut -c1-200
