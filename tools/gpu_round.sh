#!/bin/bash
# one GPU-box visit: parity tests and the default bench line — every step under a tight timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 240 python bench.py --extras-timeout 100 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.log; echo "bench rc=$?" >> gpurun_out/bench_default.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_default.log; head -c 300 gpurun_out/bench_default.json
