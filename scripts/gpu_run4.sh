cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
D="python trt-asr-engine_b200/tools/gpu_debug.py"
( for mask in 64 128 256 512 1024 2048 4 16; do echo "== TC_MASK=$mask"; PARAKEET_B200_TC_MASK=$mask PREC=1 BACKEND=2 CHUNKS=1 timeout 300 $D encoder 2>&1 | grep chunk; done ) > gpurun_out/run4.log 2>&1
tail -60 gpurun_out/run4.log
