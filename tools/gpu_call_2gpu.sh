#!/bin/bash
# 2-GPU call: N>1 parity on GPUs (SyncBN through NVLink peer memory and through NCCL, DDP wrapping) + N=2 bench A/B.
mkdir -p gpurun_out
O=gpurun_out/r2m4
nvidia-smi topo -m > ${O}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -rA -s --timeout 800 -p no:cacheprovider > ${O}_tests.txt 2>&1
echo "rc=$?" >> ${O}_tests.txt
for peer in 1; do
FSNET_BUCKETED_ALLREDUCE=$peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_n2_peer$peer.txt 2>&1
done
grep -E "passed|failed|2-GPU|DDP-wrapped|FAILED|Error" ${O}_tests.txt | cut -c1-300 | head -20
grep -E "SyncBN statistics" ${O}_bench_n2_peer*.txt | head
grep -o '"ms_per_step": [0-9.]*' ${O}_bench_n2_peer*.txt
