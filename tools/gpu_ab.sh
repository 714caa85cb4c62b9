#!/bin/bash
# quick A/B of an environment knob on the device-resident step (stereo + mono), then the parity tests
# usage: bash tools/gpu_ab.sh <tag> <KNOB> <values...>
TAG=$1; KNOB=$2; shift 2
mkdir -p gpurun_out
( timeout 300 python tools/sweep_env.py $KNOB stereo "$@"; timeout 300 python tools/sweep_env.py $KNOB mono "$@" ) > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
