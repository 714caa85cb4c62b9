#!/bin/bash
# full parity tests + sanitizer cases.  usage: gpu_t1.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
bash tools/gpu_sanitize.sh $TAG
