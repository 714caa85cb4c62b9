#!/bin/bash
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-fma-alt --no-other-scaling --mode mono"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_ -s 6 -c 2 -f -o gpurun_out/${TAG}_mono $B --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
