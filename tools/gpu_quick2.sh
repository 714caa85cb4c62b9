#!/bin/bash
# tests (with durations) + stereo/mono bench lines.  usage: gpu_quick2.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -16 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --mode mono --no-cpu --no-other-scaling > gpurun_out/${TAG}_bench_mono.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("bench","bench_mono"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        r=d["roofline"]; s=d.get("strong") or {}
        print(f, "value %.0f G  ms/step %.4f  kernel %.4f ms (alone %.4f) frac %.4f fp32 %.3f | deemph %.4f ms | strong %s | fma %.4f | e2e %.1f G frac %s | parity %s" % (
            d["value"]/1e3, d["ms_per_step"], r["kernel_ms"], r["kernel_ms_alone"], r["frac"], r["fp32_pipe"]["frac"], d["roofline_kernels"][1]["kernel_ms"],
            s.get("ms_per_step"), d["precision_fma"]["ms_per_step"], d["e2e"]["value"]/1e3, d["e2e"].get("copy_ceiling_frac"), d["parity"]["all"]))
    except Exception as e: print(f, "failed", e)
PY
