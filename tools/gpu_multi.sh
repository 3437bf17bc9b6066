#!/bin/bash
# multi-GPU visit (N = $1 GPUs): sharded-sweep tests, the driver's N-GPU bench command (replicas + sharded probe), config 4 (10M cells, sharded)
N=${1:-2}; WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt 2>&1
if [ "$WHAT" = "all" ] || [ "$WHAT" = "tests" ]; then
timeout 240 python -m pytest tests/test_sweep_shards.py tests/test_sharded_gloo.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_$N.log
tail -3 gpurun_out/pytest_multi_$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$WHAT" = "all" ] || [ "$WHAT" = "default" ]; then
timeout 420 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_default.json 2> gpurun_out/bench_${N}gpu_default.log; echo "rc=$?" >> gpurun_out/bench_${N}gpu_default.log
tail -2 gpurun_out/bench_${N}gpu_default.log; head -c 250 gpurun_out/bench_${N}gpu_default.json; echo
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "10m" ]; then
timeout 420 $TR bench.py --gpus $N --cells 10000000 --steps 2 --warmup 1 --hiters 10 > gpurun_out/bench_${N}gpu_10M_shards.json 2> gpurun_out/bench_${N}gpu_10M_shards.log; echo "rc=$?" >> gpurun_out/bench_${N}gpu_10M_shards.log
tail -2 gpurun_out/bench_${N}gpu_10M_shards.log; head -c 250 gpurun_out/bench_${N}gpu_10M_shards.json; echo
fi
