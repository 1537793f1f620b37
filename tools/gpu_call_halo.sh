#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c26
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/bench_conv.py 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.txt 2>&1; echo "$(tail -1 ${O}_bench.txt | cut -c1-200)"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > ${O}_tests_all.txt; tail -3 ${O}_tests_all.txt
