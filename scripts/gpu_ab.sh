# A/B of an environment toggle on the default bench workload: VAR=name, values "0 1 0 1"
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for E in ${VALS:-0 1 0 1}; do
env $VAR=$E timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/ab.json 2> gpurun_out/ab.err; rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print("$VAR=$E rc=$rc", round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['roofline_hbm'][0]['achieved']), d['clocks']['sm_mhz'])
PY
done
