# round 2: 256-bit epilogue stores: GEMM tests + parity, bench; ncu --set full of the streaming LayerNorm / dwconv4
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_parity_set.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -3 gpurun_out/r2p_pytest.log
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r2p_bench_new.json 2> gpurun_out/r2p_bench_new.err
PARAKEET_B200_PAIR_MODES=80 timeout 600 python bench.py $BB > gpurun_out/r2p_bench_pm80.json 2> gpurun_out/r2p_bench_pm80.err
PARAKEET_B200_PAIR_MODES=208 timeout 600 python bench.py $BB > gpurun_out/r2p_bench_pm208.json 2> gpurun_out/r2p_bench_pm208.err
PARAKEET_B200_LN_STREAM=0 PARAKEET_B200_DWCONV4=0 timeout 600 python bench.py $BB > gpurun_out/r2p_bench_lnold.json 2> gpurun_out/r2p_bench_lnold.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm|dwconv" -s 8000 -c 6 -o gpurun_out/r02_ln_stream -f python bench.py $BA > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"gemm_tc" -s 19200 -c 40 --csv --log-file gpurun_out/r2p_gemm_launches.csv python bench.py $BA > gpurun_out/r2p_ncu2.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2p_gemm_launches.csv
