#!/bin/bash
# multi-GPU scaling of the default arm at N = $NGPU
mkdir -p gpurun_out
O=gpurun_out/r2s4
N=${NGPU:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${O}_n${N}.txt 2>&1
tail -1 ${O}_n${N}.txt | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],3),'clocks',d['clocks'])"
