#!/bin/bash
mkdir -p gpurun_out
T="tests/test_terrain_post_parity.py::test_run_post_processing_default_sliders"
timeout 60 python -m pytest "$T" -m gpu -x -q > gpurun_out/diag_plain.log 2>&1; echo "rc=$?" >> gpurun_out/diag_plain.log
PB_TRACE=1 timeout 90 python -m pytest "$T" -m gpu -x -q -s > gpurun_out/diag_trace.log 2>&1; echo "rc=$?" >> gpurun_out/diag_trace.log
tail -c 600 gpurun_out/diag_plain.log; echo ----; tail -c 1500 gpurun_out/diag_trace.log
