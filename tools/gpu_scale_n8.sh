#!/bin/bash
# short 8-GPU visit (gpurun --gpus 8): multi-GPU parity tests, bench at 1 and 8 GPUs (weak + strong + per-rank parity + e2e).
# usage: bash tools/gpu_scale_n8.sh <tag>
TAG=${1:-r04scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_multi.log; tail -3 gpurun_out/${TAG}_pytest_multi.log
for N in 1 8; do
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
    e=d["e2e"]; s=d.get("strong") or {}
    print("N=$N weak %.0f G (%.4f ms)  strong %.0f G (%.4f ms, %d/GPU)  e2e %.1f G (ceiling frac %s)  parity %s / e2e %s" % (
        d["value"]/1e3, d["ms_per_step"], s.get("value",0)/1e3, s.get("ms_per_step",0), s.get("streams_per_gpu",0), e["value"]/1e3,
        e.get("copy_ceiling_frac"), d["parity"]["per_rank"], e["parity_per_shard"]))
except Exception as ex:
    print("N=$N failed", ex); import subprocess; print(subprocess.run(["tail","-5","gpurun_out/${TAG}_n$N.err"],capture_output=True,text=True).stdout)
PY
done
