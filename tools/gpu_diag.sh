#!/bin/bash
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_terrain_post_parity.py tests/test_elevation_parity.py tests/test_climate_parity.py tests/test_plates_parity.py -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_tests.log
timeout 50 python bench.py --steps 5 --warmup 2 --no-cpu --extras-timeout 30 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.log; echo "rc=$?" >> gpurun_out/final_bench.log
tail -2 gpurun_out/final_tests.log; python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read()); print(d['ms_per_step'], d['ms_steps'], d['e2e']['ms_per_step'], d['e2e']['matches_device_path'], d['throughput_in_flight'] and d['throughput_in_flight']['value'], d.get('supplementary'))"
