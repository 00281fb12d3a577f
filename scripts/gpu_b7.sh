cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x 2>&1 | tail -5
for E in 0 1 0 1; do
PARAKEET_B200_DWCONV2=$E timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/b7_$E.json 2> gpurun_out/b7_$E.err; echo "dwconv2=$E rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/b7_$E.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), d['clocks']['sm_mhz'])
PY
done
