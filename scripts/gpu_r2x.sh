# round 2: persistent decode A/B on the config-5 shape; small-batch A/B (pair kernel forced)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -3 gpurun_out/r2x_pytest.log
BB="--no-cpu-baseline --no-config3 --steps 5"
timeout 900 python bench.py $BB > gpurun_out/r2x_bench_persist.json 2> gpurun_out/r2x_bench_persist.err
PARAKEET_B200_DECODE_PERSIST=0 timeout 900 python bench.py $BB > gpurun_out/r2x_bench_graph.json 2> gpurun_out/r2x_bench_graph.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), d.get('config5_longform'), d.get('config1_offline_10s'), d.get('latency_1stream'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
bash scripts/gpu_r2w.sh
