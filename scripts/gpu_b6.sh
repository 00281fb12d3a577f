cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for E in 0 1 0 1; do
PARAKEET_B200_ATTN_EVICT=$E timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/b6_$E.json 2> gpurun_out/b6_$E.err; echo "evict=$E rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/b6_$E.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), round(d['roofline_hbm'][0]['achieved']), d['clocks']['sm_mhz'])
PY
done
