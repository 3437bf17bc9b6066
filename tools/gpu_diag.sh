#!/bin/bash
mkdir -p gpurun_out
PB_ORDERED_FLOW=1 timeout 90 python bench.py --no-cpu --steps 2 --warmup 1 --extras-timeout 60 > gpurun_out/inflight_A.json 2> gpurun_out/inflight_A.log; echo "rc=$?" >> gpurun_out/inflight_A.log
PB_BACKOFF_NS=0 timeout 90 python bench.py --no-cpu --steps 2 --warmup 1 --extras-timeout 60 > gpurun_out/inflight_B.json 2> gpurun_out/inflight_B.log; echo "rc=$?" >> gpurun_out/inflight_B.log
tail -1 gpurun_out/inflight_A.log; python -c "
import json
for f in 'AB':
    try:
        d=json.loads(open('gpurun_out/inflight_%s.json'%f).read()); print(f, d['ms_per_step'], d['throughput_in_flight'], d.get('supplementary'))
    except Exception as e: print(f, 'ERR', e)
"; tail -1 gpurun_out/inflight_B.log
