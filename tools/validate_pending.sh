#!/bin/bash
# One gpurun call that validates everything written after round 1's GPU minutes were spent, then re-measures.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/validate_pending.sh'   (~35 GPU-minutes)
# Outputs land in gpurun_out/ (merged back): pending_tests.txt, bench_default.txt, bench_prefetch.txt, gpu_tests.txt
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
FSNET_PENDING_GPU=1 timeout 1200 python -m pytest tests/test_pending_gpu.py -q -m gpu -rA --timeout 900 > gpurun_out/pending_tests.txt 2>&1
echo "pending rc=$?" >> gpurun_out/pending_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --prefetch 1 --no-cpu-baseline > gpurun_out/bench_prefetch.txt 2>&1
for w in 1 2; do   # wave-quantisation A/B of the convolution planner (numerics unchanged: same kernel, narrower channel tiles)
  FSNET_CONV_WAVE=$w timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_wave$w.txt 2>&1
  FSNET_CONV_WAVE=$w timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu > gpurun_out/conv_tests_wave$w.txt 2>&1
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.txt 2>&1
echo "gpu rc=$?" >> gpurun_out/gpu_tests.txt
tail -3 gpurun_out/smoke.txt gpurun_out/pending_tests.txt gpurun_out/gpu_tests.txt
grep -o '"e2e": {[^}]*}' gpurun_out/bench_default.txt gpurun_out/bench_prefetch.txt
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_default.txt gpurun_out/bench_wave1.txt gpurun_out/bench_wave2.txt
