#!/bin/bash
# ncu visit: launch list of one full step + full captures of the kernels the bench names
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --in-flight 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 10400 --csv --log-file gpurun_out/r02_launches_full_1M.csv $B > gpurun_out/ncu_launches.json 2> gpurun_out/ncu_launches.log; echo "rc=$?" >> gpurun_out/ncu_launches.log
P="python bench.py --workload post --steps 1 --warmup 1 --no-cpu --in-flight 0"
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SolveK -s 60 -c 2 -o gpurun_out/r02_solve -f $P > gpurun_out/ncu_solve.out 2> gpurun_out/ncu_solve.log; echo "rc=$?" >> gpurun_out/ncu_solve.log
C="python bench.py --workload climate --cells 10000000 --steps 1 --warmup 1 --no-cpu"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"SmoothFieldK|ShadowSweepK" -s 1400 -c 4 -o gpurun_out/r02_sweeps_10M -f $C > gpurun_out/ncu_sweeps.out 2> gpurun_out/ncu_sweeps.log; echo "rc=$?" >> gpurun_out/ncu_sweeps.log
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/r02_launches_full_1M.csv; grep "==" gpurun_out/ncu_solve.out gpurun_out/ncu_sweeps.out | tail -6
