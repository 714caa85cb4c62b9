#!/bin/bash
# 1/2/4/8-GPU weak-scaling bench on one box (run under gpurun --gpus 8).  usage: bash tools/gpu_scale.sh <tag>
TAG=${1:-scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2>/dev/null
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.0f Msamples/s  ms/step %.4f  e2e %.0f  frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("N=$N failed", e)
PY
done
