mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "deemph", round(d["roofline"]["deemph_kernel_ms"],4))'
for r in 0 1 2 5; do
FMB_BENCH_DATA_RANK=$r python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu --no-fma-alt 2>/dev/null | python -c "$show" "single GPU, data of rank $r"
done
FMB_BENCH_DATA_RANK=1 python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu --no-fma-alt --mode mono 2>/dev/null | python -c "$show" "mono, data of rank 1"
