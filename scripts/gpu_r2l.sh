cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BA="--no-cpu-baseline --no-latency --no-config3 --longform 0"
for pm in 16 80 144 208 16 208; do
  PARAKEET_B200_PAIR_MODES=$pm timeout 600 python bench.py $BA > gpurun_out/r2l_bench_pm$pm.json 2> gpurun_out/r2l_bench_pm$pm.err
  python - $pm <<'PY'
import json,sys
pm=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2l_bench_pm{pm}.json').read().strip().splitlines()[-1])
    print('pair modes', pm, 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
except Exception as e: print(pm,'ERR',e)
PY
done
PARAKEET_B200_PAIR_MODES=208 timeout 900 python -m pytest tests/test_gpu_parity_set.py tests/test_gpu_full_size.py tests/test_gpu_saturated_parity.py -m gpu -q -x > gpurun_out/r2l_pytest_pair.log 2>&1; tail -2 gpurun_out/r2l_pytest_pair.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_all.log; tail -5 gpurun_out/r2l_pytest_all.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2l_smoke.log 2>&1; tail -2 gpurun_out/r2l_smoke.log
