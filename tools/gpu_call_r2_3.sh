#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c3
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_model_gpu.py -q -m gpu -x --timeout 600 -p no:cacheprovider > ${O}_tests.txt 2>&1
echo "rc=$?" >> ${O}_tests.txt
timeout 200 python tools/bench_loss.py > ${O}_loss_occ8.json 2>&1
FSNET_B200_LIB=$PWD/fsnet_b200/lib/libfsnet_b200_occ6.so timeout 200 python tools/bench_loss.py > ${O}_loss_occ6.json 2>&1
FSNET_B200_LIB=$PWD/fsnet_b200/lib/libfsnet_b200_occ10.so timeout 200 python tools/bench_loss.py > ${O}_loss_occ10.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_pair -s 8 -c 1 -o ${O}_pair python tools/bench_loss.py > ${O}_ncu.log 2>&1
tail -3 ${O}_tests.txt
grep -E "fused_s" ${O}_loss_*.json
