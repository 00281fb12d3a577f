cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_cli.py tests/test_gpu_frontend.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/run36.log 2>&1
cat gpurun_out/run36.log
