# final verification of the tree as committed: full GPU suite, smoke(), default bench
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/final_tests.log; tail -6 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['latency_1stream'], d['cpu_baseline']['value'], d['clocks'])
PY
