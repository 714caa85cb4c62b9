#!/bin/bash
# A/B of libfmb variants on the mono preset + the mono-touching GPU tests under each variant.  usage: gpu_ws3.sh <tag> <mode> <variant>...
TAG=$1; MODE=$2; shift 2
mkdir -p gpurun_out
{
for n in "$@"; do
  echo "== variant $n"
  if [ "$n" = "main" ]; then L=""; else L="$PWD/tools/variants/libfmb_$n.so"; fi
  FMB_LIB_PATH=$L timeout 200 python tools/sweep_env.py FMB_PDL $MODE 1 1
  if [ "$n" != "main" ]; then
    FMB_LIB_PATH=$L timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${KSEL:-mono or silence or ragged or back_to_back or carried}" 2>&1 | tail -3
  fi
done
} > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
