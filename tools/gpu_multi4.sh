#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest "tests/test_sweep_shards.py::test_sharded_climate_peer_memory" -m gpu -x -q -k "4" > gpurun_out/pytest_multi_4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_4.log
tail -2 gpurun_out/pytest_multi_4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus 4 --steps 3 --warmup 3 --extras-timeout 90 > gpurun_out/bench_4gpu_default.json 2> gpurun_out/bench_4gpu_default.log; echo "rc=$?" >> gpurun_out/bench_4gpu_default.log
tail -1 gpurun_out/bench_4gpu_default.log
