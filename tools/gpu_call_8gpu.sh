#!/bin/bash
# 8-GPU call: scaling of the default arm and A/B of the gradient exchange (bucketed during backward vs one flat all-reduce after it)
mkdir -p gpurun_out
O=gpurun_out/r2s2
N=${NGPU:-8}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${O}_n${N}_$name.txt 2>&1
}
run default FSNET_X=1
run flat_allreduce FSNET_BUCKETED_ALLREDUCE=0
grep -E "SyncBN statistics" ${O}_n${N}_*.txt | head -3
grep -o '"ms_per_step": [0-9.]*' ${O}_n${N}_*.txt
grep -o '"value": [0-9.]*' ${O}_n${N}_*.txt | head -8
