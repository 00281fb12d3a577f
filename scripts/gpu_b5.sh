cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b50.json 2> gpurun_out/b50.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open('gpurun_out/b50.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
PY
timeout 600 python -m pytest tests/test_gpu_parity_set.py -m gpu -q -x 2>&1 | tail -3
