#!/bin/bash
TAG=$1; MODE=$2; shift 2
mkdir -p gpurun_out
{
for n in "$@"; do
  echo "== variant $n"
  if [ "$n" = "main" ]; then timeout 200 python tools/sweep_env.py FMB_PDL $MODE 1 1
  else FMB_LIB_PATH=$PWD/tools/variants/libfmb_$n.so timeout 200 python tools/sweep_env.py FMB_PDL $MODE 1 1; fi
done
} > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
