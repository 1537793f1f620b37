#!/bin/bash
# round-2 evidence: other workloads, launch list of one step, ncu captures of the two roofline kernels, smoke
mkdir -p gpurun_out
O=gpurun_out/r2p1
python __graft_entry__.py smoke 2>&1 | tail -1
for wl in cfg2b cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline > ${O}_bench_$wl.txt 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("${O}_bench_$wl.txt").read().strip().splitlines()[-1])
    print("$wl", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "enc", d["roofline_conv_encoder"].get("us_per_step"), "conv frac", d["roofline_conv"].get("frac"))
except Exception as e:
    print("$wl failed", e); print(open("${O}_bench_$wl.txt").read()[-1500:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_pair_kernel -s 1 -c 1 -o ${O}_pair python tools/bench_loss.py > ${O}_ncu_pair.log 2>&1; tail -2 ${O}_ncu_pair.log
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 1 -o ${O}_halo python tools/bench_conv.py l1_64_48x160 > ${O}_ncu_halo.log 2>&1; tail -1 ${O}_ncu_halo.log
