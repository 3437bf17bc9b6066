#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_sweep_shards.py -m gpu -x -q > gpurun_out/pytest_multi_2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_2.log
tail -2 gpurun_out/pytest_multi_2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.log; echo "rc=$?" >> gpurun_out/bench_2gpu_default.log
tail -1 gpurun_out/bench_2gpu_default.log
