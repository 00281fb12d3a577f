cd ${GRAFT_REPO_ROOT:-/root/repo}/trt-asr-engine_b200
K=bin/kbench
echo "== cold weights (rot) vs warm"
$K gemm 6144 4096 1024 128 silu 0 1 | tail -1
$K gemm 6144 4096 1024 128 silu 0 32 | tail -1
$K gemm 6144 1024 4096 128 partial2pb 0 1 | tail -1
$K gemm 6144 1024 4096 128 partial2pb 0 32 | tail -1
$K gemm 6144 1024 1024 256 partial1b 0 1 | tail -1
$K gemm 6144 1024 1024 256 partial1b 0 128 | tail -1
$K gemm 6144 3072 1024 128 f32 0 1 | tail -1
$K gemm 6144 3072 1024 128 f32 0 32 | tail -1
echo "== with LN interleaved (L2 thrash by x)"
$K gemm 6144 4096 1024 128 silu 1 32 | tail -1
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
