cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python bench.py ${BARGS:-} > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_last.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d.get('latency_1stream'), d['clocks'])
PY
