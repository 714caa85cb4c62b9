#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "whole_stream_runs or config_3" > gpurun_out/${TAG}_first.log 2>&1; echo "first exit $?" >> gpurun_out/${TAG}_first.log; tail -15 gpurun_out/${TAG}_first.log
if grep -q "first exit 0" gpurun_out/${TAG}_first.log; then
  timeout -k 5 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
  ( FMB_WS=0 timeout 200 python tools/sweep_env.py FMB_PDL stereo 1; FMB_WS=1 timeout 200 python tools/sweep_env.py FMB_PDL stereo 1 1 ) > gpurun_out/${TAG}_ab.txt 2>&1; cat gpurun_out/${TAG}_ab.txt
fi
