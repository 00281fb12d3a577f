cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
  timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
) > gpurun_out/full.log 2>&1
tail -12 gpurun_out/full.log
