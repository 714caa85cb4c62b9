#!/bin/bash
# sweep of the CTA phase stagger (clocks) of the demod kernel.  usage: bash tools/gpu_stagger.sh <tag> <mode> v1 v2 ...
TAG=$1; MODE=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  FMB_STAGGER=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --mode $MODE --no-e2e 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stagger', $v, 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))" | tee -a gpurun_out/${TAG}_stagger_${MODE}.txt
done
