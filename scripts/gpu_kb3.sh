cd ${GRAFT_REPO_ROOT:-/root/repo}
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_offline_long.py::test_two_utterances_ragged -m gpu -q -x 2>&1 | tail -4
cd trt-asr-engine_b200
K=bin/kbench
for D in 0 1; do
echo "== EPI_DIRECT=$D"
PARAKEET_B200_EPI_DIRECT=$D $K gemm 6144 4096 1024 128 silu | tail -1
PARAKEET_B200_EPI_DIRECT=$D $K gemm 6144 1024 1024 256 partial1b | tail -1
PARAKEET_B200_EPI_DIRECT=$D PARAKEET_B200_GEMM_BN=256 $K gemm 6144 1024 1024 256 silu | tail -1
PARAKEET_B200_EPI_DIRECT=$D $K gemm 6144 2560 128 256 silu | tail -1
done
