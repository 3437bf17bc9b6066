#!/bin/bash
mkdir -p gpurun_out
for m in none smi nvml; do
BENCH_SAMPLER=$m timeout 70 python bench.py --steps 6 --warmup 2 --no-cpu --in-flight 0 > gpurun_out/sampler_$m.json 2> gpurun_out/sampler_$m.log; echo "rc=$?" >> gpurun_out/sampler_$m.log
python -c "
import json
d=json.loads(open('gpurun_out/sampler_$m.json').read()); print('$m', d['ms_per_step'], d['ms_steps'], d['clocks'])"
done
