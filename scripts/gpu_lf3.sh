cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/lf_tests3.log
tail -8 gpurun_out/lf_tests3.log
PARAKEET_B200_LF_ATTN_TMA=0 timeout 300 python scripts/lf_probe.py 1200 2 2>&1 | tail -1
timeout 300 python scripts/lf_probe.py 1200 2 2>&1 | tail -1
timeout 300 python scripts/lf_probe.py 3600 2 2>&1 | tail -1
