#!/bin/bash
# cluster-multicast A/B of the convolution kernel: micro-benchmarks, parity tests and per-layer timings per cluster size
mkdir -p gpurun_out
O=gpurun_out/r2c12
(cd tools/ubench && timeout 60 ./umma_shift; timeout 120 ./mc_bw) > ${O}_ubench.txt 2>&1
for mc in 1 2 4 8; do
  FSNET_CONV_MC=$mc FSNET_CONV_MC_PRINT=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -4 > ${O}_tests_mc$mc.txt
  echo "MC=$mc: $(tail -1 ${O}_tests_mc$mc.txt)"
  FSNET_CONV_MC=$mc timeout 200 python tools/bench_conv.py l1_64_48x160 l2_128_24x80 l3_256_12x40 l4_512_6x20 dec96_96x320 2>&1 | tail -1
done
cat ${O}_ubench.txt
