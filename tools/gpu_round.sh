#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, one full ncu capture of the demod kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --mode mono --no-cpu > gpurun_out/${TAG}_bench_mono.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_mono.json
timeout 300 python bench.py --steps 20 --warmup 3 --precision fma --no-cpu > gpurun_out/${TAG}_bench_fma.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_fma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu --no-fma-alt > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 2 -f -o gpurun_out/${TAG}_demod \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 1 -f -o gpurun_out/${TAG}_demod_mono \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt --mode mono >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_deemph -s 3 -c 1 -f -o gpurun_out/${TAG}_deemph \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt >> gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
