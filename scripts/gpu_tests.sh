cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/tests.log
tail -40 gpurun_out/tests.log
