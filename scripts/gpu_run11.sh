cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( PRE=g1 PREC=1 BACKEND=1 timeout 300 $D saturated
  PRE=g0 PREC=1 BACKEND=1 timeout 300 $D saturated
  PRE=f1 PREC=1 BACKEND=1 timeout 300 $D saturated
  PRE=f1,g1,g0 PREC=1 BACKEND=1 timeout 300 $D saturated
) > gpurun_out/run11.log 2>&1
grep -E "trial 0|trial 3|pre-engine|Error" gpurun_out/run11.log | tail -40
