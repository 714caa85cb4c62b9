#!/bin/bash
# one ncu --set full capture of the stereo demod kernel of the default build.  usage: gpu_ncu2.sh <tag>
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-fma-alt --no-other-scaling"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 1 -f -o gpurun_out/${TAG}_demod $B --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200; ls -la gpurun_out | grep ${TAG}
