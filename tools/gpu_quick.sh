#!/bin/bash
# quick GPU visit: parity tests + bench (no ncu).  usage: bash tools/gpu_quick.sh <tag> [pytest-k-filter]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --mode mono --no-cpu > gpurun_out/${TAG}_bench_mono.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_mono.json
