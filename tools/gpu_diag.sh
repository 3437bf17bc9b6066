#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_climate_parity.py tests/test_large_parity.py -m gpu -x -q > gpurun_out/diag_tests.log 2>&1; echo "rc=$?" >> gpurun_out/diag_tests.log
timeout 150 python bench.py --workload climate --cells 10000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ab_land.json 2> gpurun_out/ab_land.log; echo "rc=$?" >> gpurun_out/ab_land.log
timeout 100 python bench.py --workload climate --steps 3 --warmup 2 --no-cpu > gpurun_out/ab_land_1M.json 2> gpurun_out/ab_land_1M.log; echo "rc=$?" >> gpurun_out/ab_land_1M.log
tail -3 gpurun_out/diag_tests.log; grep -A8 "per-kernel device time" gpurun_out/ab_land.log; head -c 200 gpurun_out/ab_land.json; echo; grep -A8 "per-kernel device time" gpurun_out/ab_land_1M.log; head -c 200 gpurun_out/ab_land_1M.json
