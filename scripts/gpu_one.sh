# run the pytest selection given in $SEL on the GPU box
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1200 python -m pytest ${SEL:-tests} -m gpu -q -x -s 2>&1 | tail -25 > gpurun_out/one.log
tail -25 gpurun_out/one.log
