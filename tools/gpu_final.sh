#!/bin/bash
# final 1-GPU visit of a round: tests, smoke, bench lines (stereo with CPU baseline, mono, fma), general ratios,
# raw copy ceiling, ncu launch list + full captures.  usage: gpu_final.sh <tag>
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1; nproc >> gpurun_out/${TAG}_smi.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --mode mono --no-cpu --no-other-scaling > gpurun_out/${TAG}_bench_mono.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --precision fma --no-cpu --no-other-scaling > gpurun_out/${TAG}_bench_fma.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --scaling strong --no-cpu --no-fma-alt > gpurun_out/${TAG}_bench_strong.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("reference_arm","bench","bench_mono","bench_fma","bench_strong"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        if f=="reference_arm": print(f, d["value"], d["cpu_baseline"]["cores"], d.get("maps_product_library")); continue
        r=d["roofline"]; e=d["e2e"]
        print(f, "value %.0f G  ms/step %.4f  kernel %s %.4f ms (alone %.4f) frac %.4f fp32 %.3f | deemph %.4f ms frac %.3f | e2e %.1f G ceil %s | parity %s | cpu %s" % (
            d["value"]/1e3, d["ms_per_step"], r["kernel"], r["kernel_ms"], r["kernel_ms_alone"], r["frac"], r["fp32_pipe"]["frac"], d["roofline_kernels"][1]["kernel_ms"], d["roofline_kernels"][1]["frac"],
            e["value"]/1e3, e.get("copy_ceiling_frac"), d["parity"]["all"], (d.get("cpu_baseline") or {}).get("value")))
    except Exception as ex: print(f, "failed", ex)
PY
{ for cfg in "240000 48000 2 90 0" "240000 48000 1 128 0" "250000 44100 2 90 0" "192000 48000 0 90 0" "192000 48000 2 128 0" "192000 48000 2 90 1"; do timeout 120 python tools/time_config.py $cfg; done; } > gpurun_out/${TAG}_general_ratios.txt 2>&1; cat gpurun_out/${TAG}_general_ratios.txt
timeout 600 python tools/h2d_ceiling.py > gpurun_out/${TAG}_h2d_ceiling.txt 2>&1; cat gpurun_out/${TAG}_h2d_ceiling.txt
B="python bench.py --no-cpu --no-fma-alt --no-other-scaling"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $B --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fmb_demod -s 3 -c 2 -f -o gpurun_out/${TAG}_demod $B --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_mono_ws -s 3 -c 1 -f -o gpurun_out/${TAG}_mono_ws $B --steps 3 --warmup 3 --mode mono >> gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_deemph -s 3 -c 1 -f -o gpurun_out/${TAG}_deemph $B --steps 3 --warmup 3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG} | tail -30
