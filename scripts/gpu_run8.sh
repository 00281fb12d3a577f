cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( PREC=1 BACKEND=1 timeout 300 $D saturated
  PREC=1 BACKEND=0 timeout 300 $D saturated
  PREC=1 BACKEND=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 $D saturated 2>&1 | tail -40
) > gpurun_out/run8.log 2>&1
tail -70 gpurun_out/run8.log
