#!/bin/bash
# full GPU test suite on the default build, then step-time A/B of the default build against variants.
# usage: gpu_ab2.sh <tag> <mode> <variant>...
TAG=$1; MODE=$2; shift 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -5 gpurun_out/${TAG}_pytest.log
{
echo "== default build"; timeout 200 python tools/sweep_env.py FMB_PDL $MODE 1 1
for n in "$@"; do echo "== variant $n"; FMB_LIB_PATH=$PWD/tools/variants/libfmb_$n.so timeout 200 python tools/sweep_env.py FMB_PDL $MODE 1 1; done
} > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
