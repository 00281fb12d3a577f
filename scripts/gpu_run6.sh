cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"; cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
  LAYERS=24 PREC=1 BACKEND=0 STREAMS=8 ROWS=64 NS=6 CHUNKS=6 timeout 900 $D stream 2>&1 | grep -v "step "
  LAYERS=24 PREC=0 BACKEND=0 STREAMS=8 ROWS=64 NS=6 CHUNKS=6 timeout 900 $D stream 2>&1 | grep -v "step "
  LAYERS=24 PREC=0 BACKEND=0 CHUNKS=3 timeout 600 $D encoder
  LAYERS=24 PREC=1 BACKEND=0 CHUNKS=3 timeout 600 $D encoder
) > gpurun_out/run6.log 2>&1
tail -60 gpurun_out/run6.log
