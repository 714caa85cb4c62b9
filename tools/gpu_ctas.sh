#!/bin/bash
TAG=$1
mkdir -p gpurun_out
{
echo "== fmb_demod_kernel (FMB_WS=0), CTAs per SM 1 2 3"
FMB_WS=0 timeout 300 python tools/sweep_env.py FMB_MAX_CTAS_PER_SM stereo 1 2 3
echo "== fmb_stereo_ws_kernel, CTAs per SM 1 2"
FMB_WS=1 timeout 300 python tools/sweep_env.py FMB_MAX_CTAS_PER_SM stereo 1 2
echo "== mono: fmb_demod_kernel 1 2 3, ws 1 2"
FMB_WS=0 timeout 300 python tools/sweep_env.py FMB_MAX_CTAS_PER_SM mono 1 2 3
FMB_WS=1 timeout 300 python tools/sweep_env.py FMB_MAX_CTAS_PER_SM mono 1 2
} > gpurun_out/${TAG}_ctas.txt 2>&1
cat gpurun_out/${TAG}_ctas.txt
