cd ${GRAFT_REPO_ROOT:-/root/repo}/trt-asr-engine_b200
K=bin/kbench
( nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 250 > /tmp/smi.log & echo $! > /tmp/smi.pid )
$K gemm 6144 4096 1024 20000 silu | tail -3
$K gemm 6144 1024 4096 20000 partial2pb | tail -3
kill $(cat /tmp/smi.pid)
sort /tmp/smi.log | uniq -c | sort -k1 -n -r | head -12
