#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --workload climate --cells 10000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ab_packed.json 2> gpurun_out/ab_packed.log; echo "rc=$?" >> gpurun_out/ab_packed.log
PB_NO_PACKED_ROWS=1 timeout 150 python bench.py --workload climate --cells 10000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ab_csr.json 2> gpurun_out/ab_csr.log; echo "rc=$?" >> gpurun_out/ab_csr.log
grep -A12 "per-kernel device time" gpurun_out/ab_packed.log | head -14; grep -A12 "per-kernel device time" gpurun_out/ab_csr.log | head -14
