#!/bin/bash
# last GPU-box visit of round 2 (≈ 100 s of budget left): the whole -m gpu suite without -x, then the export probe, then one
# ncu capture of the two export kernels — each step under its own timeout, most important first
mkdir -p gpurun_out
timeout 62 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log
timeout 22 python tools/export_probe.py 1000000 4096 3 > gpurun_out/r02_export_probe.json 2> gpurun_out/r02_export_probe.err; echo "probe rc=$?"
head -c 1500 gpurun_out/r02_export_probe.json; tail -2 gpurun_out/r02_export_probe.err
timeout 28 ncu --set full --clock-control none --import-source on -k regex:Map -c 2 -o gpurun_out/r02_export_ncu -f python tools/export_probe.py 250000 2048 1 > gpurun_out/r02_export_ncu.log 2>&1; echo "ncu rc=$?"
