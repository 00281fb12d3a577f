cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( PREC=1 BACKEND=1 timeout 600 $D lensweep ) > gpurun_out/run12.log 2>&1
tail -60 gpurun_out/run12.log
