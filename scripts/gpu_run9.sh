cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for k in "precise-simt and (closed_loop or functional or saturated)" "precise-simt and (closed_loop or saturated)" "precise-simt and (functional or saturated)" "precise-auto and (functional or saturated)"; do
    echo "=== -k $k"; timeout 600 python -m pytest tests/test_gpu_model.py -q -k "$k" 2>&1 | grep -E "passed|failed|AssertionError:" ; done
) > gpurun_out/run9.log 2>&1
tail -30 gpurun_out/run9.log
