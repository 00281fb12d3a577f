cd ${GRAFT_REPO_ROOT:-/root/repo}/trt-asr-engine_b200
K=bin/kbench
echo "== N=1024 K=1024 residual shapes"
$K gemm 6144 1024 1024 200 partial1b | tail -1
PARAKEET_B200_GEMM_BN=256 $K gemm 6144 1024 1024 200 partial1b | tail -1
PARAKEET_B200_GEMM_BN=512 $K gemm 6144 1024 1024 200 partial1b | tail -1
PARAKEET_B200_GEMM_BN=512 $K gemm 6144 1024 1024 200 partial2pb | tail -1
$K gemm 6144 1024 1024 200 partial2b | tail -1
echo "== FFN2 K=4096"
$K gemm 6144 1024 4096 100 partial1b | tail -1
PARAKEET_B200_GEMM_BN=512 $K gemm 6144 1024 4096 100 partial2pb | tail -1
echo "== FFN1 silu N=4096"
$K gemm 6144 4096 1024 100 silu | tail -1
PARAKEET_B200_GEMM_BN=512 $K gemm 6144 4096 1024 100 silu | tail -1
PARAKEET_B200_GEMM_BN=128 $K gemm 6144 4096 1024 100 silu | tail -1
echo "== GLU N=2048"
$K gemm 6144 2048 1024 100 glu | tail -1
PARAKEET_B200_GEMM_BN=512 $K gemm 6144 2048 1024 100 glu | tail -1
PARAKEET_B200_GEMM_BN=128 $K gemm 6144 2048 1024 100 glu | tail -1
