#!/bin/bash
# the mono warp-specialised kernel off the 4:1 fast path: full GPU tests, then step times with and without it.  usage: gpu_wsgen.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
{
for g in 0 1; do
  for cfg in "240000 48000 1 128 0" "240000 48000 1 90 0" "250000 44100 1 128 0" "240000 32000 1 128 1"; do
    echo -n "FMB_WS_GENERIC=$g  "; FMB_WS_GENERIC=$g timeout 120 python tools/time_config.py $cfg
  done
done
} > gpurun_out/${TAG}_ws_generic.txt 2>&1; cat gpurun_out/${TAG}_ws_generic.txt
