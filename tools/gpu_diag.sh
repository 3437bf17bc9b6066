#!/bin/bash
mkdir -p gpurun_out
for m in 1 0; do
BENCH_STEP_SYNC=$m timeout 60 python bench.py --steps 6 --warmup 2 --no-cpu --in-flight 0 > gpurun_out/stepsync_$m.json 2> gpurun_out/stepsync_$m.log; echo "rc=$?" >> gpurun_out/stepsync_$m.log
python -c "
import json
d=json.loads(open('gpurun_out/stepsync_$m.json').read()); print('sync=$m', d['ms_per_step'], d['ms_steps'], d['e2e']['ms_per_step'])"
done
