#!/bin/bash
# Round-2 first GPU call: parity holes (full-size parity, pending tests, log-image fix), A/B of switches written in round 1, launch list.
mkdir -p gpurun_out
O=gpurun_out/r2c1
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "smoke rc=$?" >> ${O}_smoke.txt
FSNET_PENDING_GPU=1 timeout 1500 python -m pytest tests -q -m gpu -rA -s --timeout 900 -p no:cacheprovider > ${O}_gpu_tests.txt 2>&1
echo "gpu rc=$?" >> ${O}_gpu_tests.txt
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench_default.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --prefetch 1 --no-cpu-baseline > ${O}_bench_prefetch.txt 2>&1
for w in 1 2; do
  FSNET_CONV_WAVE=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_wave$w.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file ${O}_launches.csv python tools/profile_step.py > ${O}_ncu.log 2>&1
grep -E "passed|failed|error" ${O}_gpu_tests.txt | tail -5
grep -E "^\[|FAILED|ERROR" ${O}_gpu_tests.txt | head -40
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_*.txt
grep -o '"e2e": {[^}]*}' ${O}_bench_default.txt ${O}_bench_prefetch.txt
