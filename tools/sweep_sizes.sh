mkdir -p gpurun_out
for S in 64 256 512 1024 8192; do
  echo "streams $S: default policy, forced static (FMB_CHUNK=0), forced dynamic 2:20"
  SWEEP_STREAMS=$S python tools/sweep_env.py FMB_DUMMY stereo x 2>&1 | tail -1
  SWEEP_STREAMS=$S python tools/sweep_env.py FMB_CHUNK stereo 0 2>&1 | tail -1
  SWEEP_STREAMS=$S python tools/sweep_env.py FMB_CHUNK,FMB_TAIL_PCT stereo 2:20 2>&1 | tail -1
done
