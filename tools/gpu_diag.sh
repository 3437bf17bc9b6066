#!/bin/bash
mkdir -p gpurun_out
T="tests/test_delaunator.py::test_pipeline_parity_on_delaunator_mesh"
timeout 40 python -m pytest "$T" -m gpu -x -q > gpurun_out/diag_plain.log 2>&1; echo "rc=$?" >> gpurun_out/diag_plain.log
tail -c 400 gpurun_out/diag_plain.log
