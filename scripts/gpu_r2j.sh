cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py tests/test_cli.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
for k in ('config3_64streams_mixed_cache','config5_longform','config1_offline_10s','latency_1stream','cpu_baseline'): print(k, d.get(k))
PY
