cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( PREC=1 BACKEND=0 CHUNKS=2 timeout 300 $D encoder;
  PREC=1 BACKEND=2 CHUNKS=2 timeout 300 $D encoder;
  PREC=1 BACKEND=2 timeout 300 $D decode;
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/run3.log 2>&1
tail -60 gpurun_out/run3.log
