#!/bin/bash
mkdir -p gpurun_out
timeout 240 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.log; echo "bench rc=$?" >> gpurun_out/bench_default.log
tail -2 gpurun_out/bench_default.log; head -c 300 gpurun_out/bench_default.json
