#!/bin/bash
# step-time A/B of variants on the stereo preset and on the 128-tap stereo configuration.  usage: gpu_ab3.sh <tag> <variant>...
TAG=$1; shift
mkdir -p gpurun_out
{
for n in "$@"; do
  echo "== variant $n"
  if [ "$n" = "main" ]; then L=""; else L="$PWD/tools/variants/libfmb_$n.so"; fi
  FMB_LIB_PATH=$L timeout 200 python tools/sweep_env.py FMB_PDL stereo 1 1
  FMB_LIB_PATH=$L timeout 120 python tools/time_config.py 192000 48000 2 128 0
done
} > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
