cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BA="--no-cpu-baseline --no-latency --no-config3"
timeout 600 python bench.py $BA > gpurun_out/r2k_bench_base.json 2> gpurun_out/r2k_bench_base.err
PARAKEET_B200_PAIR_MODES=16 timeout 600 python bench.py $BA > gpurun_out/r2k_bench_pair_silu.json 2> gpurun_out/r2k_bench_pair_silu.err
timeout 600 python bench.py $BA > gpurun_out/r2k_bench_base2.json 2> gpurun_out/r2k_bench_base2.err
PARAKEET_B200_PAIR_MODES=16 timeout 600 python bench.py $BA > gpurun_out/r2k_bench_pair_silu2.json 2> gpurun_out/r2k_bench_pair_silu2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity_set.py -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; tail -2 gpurun_out/r2k_pytest.log
PARAKEET_B200_PAIR_MODES=16 timeout 900 python -m pytest tests/test_gpu_parity_set.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2k_pytest_pair.log 2>&1; tail -2 gpurun_out/r2k_pytest_pair.log
timeout 300 python bench.py --no-cpu-baseline --no-config3 --steps 5 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config5_longform'])"
