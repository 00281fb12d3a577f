cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/run1.log 2>&1
( timeout 300 $D frontend; echo "rc=$?";
  PREC=1 BACKEND=1 timeout 300 $D decode; echo "rc=$?";
  PREC=1 BACKEND=1 timeout 300 $D encoder; echo "rc=$?";
  PREC=1 BACKEND=1 timeout 300 $D stream; echo "rc=$?";
  PREC=0 BACKEND=1 timeout 300 $D encoder; echo "rc=$?";
  PREC=1 timeout 120 $D gemm; echo "rc=$?";
  PREC=0 timeout 120 $D gemm; echo "rc=$?";
  PREC=1 BACKEND=0 timeout 300 $D encoder; echo "rc=$?";
  PREC=1 BACKEND=0 ROWS=256 STREAMS=32 NS=24 CHUNKS=4 timeout 300 $D stream; echo "rc=$?" ) >> gpurun_out/run1.log 2>&1
tail -80 gpurun_out/run1.log
