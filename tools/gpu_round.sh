#!/bin/bash
# one GPU-box visit: parity tests, a debug-timed step, the default bench line — every step under a tight timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
PB_DEBUG=1 timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu --in-flight 0 > gpurun_out/bench_debug.json 2> gpurun_out/bench_debug.log; echo "rc=$?" >> gpurun_out/bench_debug.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.log; echo "bench rc=$?" >> gpurun_out/bench_default.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_debug.log; tail -2 gpurun_out/bench_default.log; head -c 300 gpurun_out/bench_default.json
