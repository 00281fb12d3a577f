cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_offline_long.py::test_two_utterances_ragged -m gpu -q -rA 2>&1 | tail -8 > gpurun_out/lf_tests2.log
tail -8 gpurun_out/lf_tests2.log
timeout 900 python bench.py --longform 4 > gpurun_out/bench_lf.json 2> gpurun_out/bench_lf.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_lf.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_lf.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for r in d['roofline_hbm']: print(r['kernel'][:50], round(r['achieved'],1), round(r['frac'],3))
print(d.get('longform'))
PY
