#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_terrain_post_parity.py tests/test_large_parity.py tests/test_delaunator.py -m gpu -x -q > gpurun_out/diag_tests.log 2>&1; echo "rc=$?" >> gpurun_out/diag_tests.log
timeout 150 python bench.py --workload post --steps 3 --warmup 2 --no-cpu --in-flight 0 > gpurun_out/diag_post.json 2> gpurun_out/diag_post.log; echo "rc=$?" >> gpurun_out/diag_post.log
tail -3 gpurun_out/diag_tests.log; grep -A8 "per-kernel device time" gpurun_out/diag_post.log; head -c 200 gpurun_out/diag_post.json
