cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/tests_t4.log
tail -8 gpurun_out/tests_t4.log
