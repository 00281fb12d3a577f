cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x 2>&1 | tail -5
LF=32 bash scripts/gpu_longform.sh
