#!/bin/bash
# final check of the committed tree: smoke, full GPU suite, the default bench line
mkdir -p gpurun_out
O=gpurun_out/r2f4
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > ${O}_tests_all.txt; tail -2 ${O}_tests_all.txt
s=$(date +%s); python bench.py > ${O}_bench_default.txt 2> ${O}_bench_default.err; echo "default bench: $(( $(date +%s) - s )) s"
tail -1 ${O}_bench_default.txt | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'pair', round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],4), 'conv', round(d['roofline_conv']['frac'],4), round(d['roofline_conv_encoder']['frac_on_pipe'],3), 'cpu', round(d['cpu_baseline']['value'],2), d['gpu_reference']['cudnn_tf32'])"
