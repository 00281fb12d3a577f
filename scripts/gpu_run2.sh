cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( PREC=1 timeout 120 $D gemm; echo "rc=$?";
  for mask in 1 2 32 0; do echo "== TC_MASK=$mask"; PARAKEET_B200_TC_MASK=$mask PREC=1 BACKEND=0 CHUNKS=2 timeout 300 $D encoder; done ) > gpurun_out/run2.log 2>&1
tail -60 gpurun_out/run2.log
