#!/bin/bash
# Round-2 second GPU call: frame-pair loss kernel -- parity, timings (occupancy 10 / 8 / old kernel), one ncu --set full capture.
mkdir -p gpurun_out
O=gpurun_out/r2c2
FSNET_PENDING_GPU=1 timeout 1500 python -m pytest tests -q -m gpu -rA -s --timeout 900 -p no:cacheprovider > ${O}_gpu_tests.txt 2>&1
echo "gpu rc=$?" >> ${O}_gpu_tests.txt
timeout 200 python tools/bench_loss.py > ${O}_loss_occ10.json 2>&1
FSNET_B200_LIB=$PWD/fsnet_b200/lib/libfsnet_b200_occ8.so timeout 200 python tools/bench_loss.py > ${O}_loss_occ8.json 2>&1
FSNET_LOSS_PAIR=0 timeout 200 python tools/bench_loss.py > ${O}_loss_old.json 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench_default.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --prefetch 1 --no-cpu-baseline > ${O}_bench_prefetch.txt 2>&1
for w in 1 2; do
  FSNET_CONV_WAVE=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_wave$w.txt 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_pair -s 8 -c 2 -o ${O}_pair python tools/bench_loss.py > ${O}_ncu.log 2>&1
grep -E "passed|failed" ${O}_gpu_tests.txt | tail -3
grep -E "FAILED|ERROR|gradient tensors" ${O}_gpu_tests.txt | cut -c1-300 | head -30
grep -E "fused_s|bwd_s0" ${O}_loss_*.json
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_*.txt
grep -o '"e2e": {[^}]*}' ${O}_bench_default.txt ${O}_bench_prefetch.txt
grep -o '"roofline": {[^}]*}' ${O}_bench_default.txt
