#!/bin/bash
mkdir -p gpurun_out
timeout 240 python bench.py --extras-timeout 100 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.log; echo "bench rc=$?" >> gpurun_out/bench_default.log
python -c "
import json
d=json.loads(open('gpurun_out/bench_default.json').read()); print(d['ms_per_step'], d['ms_steps'], d['e2e']['ms_per_step'], d['clocks'])"
