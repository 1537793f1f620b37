#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c11
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x --timeout 600 -p no:cacheprovider > ${O}_tests.txt 2>&1
echo "rc=$?" >> ${O}_tests.txt
for occ in 0 1 2; do
FSNET_CONV_OCC=$occ timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_occ$occ.txt 2>&1
done
FSNET_CONV_OCC=2 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -m gpu -x --timeout 600 -p no:cacheprovider > ${O}_tests_occ2.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file ${O}_launches.csv python tools/profile_step.py > ${O}_ncu.log 2>&1
tail -2 ${O}_tests.txt ${O}_tests_occ2.txt
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_*.txt
