#!/bin/bash
# final evidence of round 2: full GPU suite, default bench line, other workloads, launch list, ncu captures of the two roofline kernels
mkdir -p gpurun_out
O=gpurun_out/r2f2
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > ${O}_tests_all.txt; tail -2 ${O}_tests_all.txt
s=$(date +%s); python bench.py > ${O}_bench_default.txt 2> ${O}_bench_default.err; echo "default bench: $(( $(date +%s) - s )) s"
for wl in cfg2b cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline > ${O}_bench_$wl.txt 2>&1
done
python - <<PY
import json
for n in ("default","cfg2b","cfg3","cfg4","cfg5"):
    try:
        d=json.loads(open("${O}_bench_%s.txt" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "; pair us", round(d["roofline"]["avg_launch_us"],1), "frac", round(d["roofline"]["frac"],4), "; conv frac", round(d["roofline_conv"]["frac"],4), "enc us", d["roofline_conv_encoder"].get("us_per_step"))
    except Exception as e:
        print(n, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_pair_kernel -s 1 -c 1 -o ${O}_pair python tools/bench_loss.py > ${O}_ncu_pair.log 2>&1; tail -1 ${O}_ncu_pair.log
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 1 -o ${O}_halo python tools/bench_conv.py l1_64_48x160 > ${O}_ncu_halo.log 2>&1; tail -1 ${O}_ncu_halo.log
