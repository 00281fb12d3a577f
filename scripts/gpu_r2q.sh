# round 2: streaming LayerNorm 16 warps x 2 rows vs 8 x 3: parity subset, bench A/B, per-kernel durations
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_pytest.log
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r2q_bench_new.json 2> gpurun_out/r2q_bench_new.err
PARAKEET_B200_LN_SMEM_KB=190 timeout 600 python bench.py $BB > gpurun_out/r2q_bench_ln8.json 2> gpurun_out/r2q_bench_ln8.err
timeout 600 python bench.py $BB > gpurun_out/r2q_bench_new2.json 2> gpurun_out/r2q_bench_new2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"layernorm|dwconv" -s 8000 -c 60 --csv --log-file gpurun_out/r2q_ln.csv python bench.py $BA > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2q_ln.csv
