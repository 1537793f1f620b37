#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2m10
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_n1.txt 2>&1; echo "n1 $(grep -o '"ms_per_step": [0-9.]*' ${O}_n1.txt | head -2 | tr '\n' ' ')"
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > ${O}_$name.txt 2>&1
  echo "$name $(grep -o '"ms_per_step": [0-9.]*' ${O}_$name.txt | head -2 | tr '\n' ' ')"
}
run flat FSNET_X=1
run bucketed FSNET_BUCKETED_ALLREDUCE=1
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_optim_gpu.py tests/test_train_script_gpu.py -x -q -m gpu 2>&1 | tail -1
