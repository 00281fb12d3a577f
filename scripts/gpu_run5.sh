cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( for i in 1 2 3; do PREC=1 BACKEND=2 CHUNKS=2 timeout 300 $D encoder 2>&1 | grep chunk; done
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/run5.log 2>&1
tail -70 gpurun_out/run5.log
