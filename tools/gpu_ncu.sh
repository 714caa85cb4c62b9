#!/bin/bash
# ncu captures: launch list + full capture of the demod kernel (stereo, mono) and the de-emphasis kernel
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-fma-alt --no-other-scaling"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $B --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 2 -f -o gpurun_out/${TAG}_demod $B --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 1 -f -o gpurun_out/${TAG}_demod_mono $B --steps 3 --warmup 3 --mode mono >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_deemph -s 3 -c 1 -f -o gpurun_out/${TAG}_deemph $B --steps 3 --warmup 3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
