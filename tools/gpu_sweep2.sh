#!/bin/bash
TAG=$1
mkdir -p gpurun_out
{
for m in stereo mono; do
  timeout 400 python tools/sweep_env.py FMB_CHUNK,FMB_TAIL_PCT $m 2:22 8:0 2:5 2:10 4:10 1:5 2:40
done
} > gpurun_out/${TAG}_sweep.txt 2>&1
cat gpurun_out/${TAG}_sweep.txt
