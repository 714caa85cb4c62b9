#!/bin/bash
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-fma-alt --no-other-scaling"
FMB_WS=0 FMB_MAX_CTAS_PER_SM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 1 -f -o gpurun_out/${TAG}_demod_1cta $B --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
