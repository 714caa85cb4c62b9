#!/bin/bash
# Round-2 GPU visit: parity tests (all, no -x), smoke, bench (stereo, mono), copy ceiling.
# usage (under gpurun): bash tools/gpu_r02.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
if [ "${2:-tests}" = "tests" ]; then
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -25 gpurun_out/${TAG}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --mode mono --no-cpu --no-other-scaling > gpurun_out/${TAG}_bench_mono.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_mono.json
timeout 600 python tools/h2d_ceiling.py > gpurun_out/${TAG}_h2d_ceiling.txt 2>&1; cat gpurun_out/${TAG}_h2d_ceiling.txt
if [ "${3:-noncu}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
     python bench.py --steps 4 --warmup 3 --no-cpu --no-fma-alt --no-other-scaling > gpurun_out/${TAG}_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 2 -f -o gpurun_out/${TAG}_demod \
     python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt --no-other-scaling > gpurun_out/${TAG}_ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 1 -f -o gpurun_out/${TAG}_demod_mono \
     python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt --no-other-scaling --mode mono >> gpurun_out/${TAG}_ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_deemph -s 3 -c 1 -f -o gpurun_out/${TAG}_deemph \
     python bench.py --steps 3 --warmup 3 --no-cpu --no-fma-alt --no-other-scaling >> gpurun_out/${TAG}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -20
