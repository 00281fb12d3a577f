# round 2: streaming LayerNorm + weights-before-wait: parity subset, then bench A/B (LN old/new, CTA-pair epilogue modes)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_parity_set.py tests/test_gpu_full_size.py tests/test_gpu_engine_api.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -4 gpurun_out/r2n_pytest.log
BA="--no-cpu-baseline --no-config3 --longform 0"
timeout 600 python bench.py $BA > gpurun_out/r2n_bench_new.json 2> gpurun_out/r2n_bench_new.err
PARAKEET_B200_LN_STREAM=0 timeout 600 python bench.py $BA --no-latency > gpurun_out/r2n_bench_lnold.json 2> gpurun_out/r2n_bench_lnold.err
for pm in 80 144 208; do
  PARAKEET_B200_PAIR_MODES=$pm timeout 600 python bench.py $BA --no-latency > gpurun_out/r2n_bench_pm$pm.json 2> gpurun_out/r2n_bench_pm$pm.err
done
timeout 600 python bench.py $BA --no-latency > gpurun_out/r2n_bench_new2.json 2> gpurun_out/r2n_bench_new2.err
for n in 128 256; do timeout 600 python bench.py --streams $n $BA --no-latency > gpurun_out/r2n_bench_$n.json 2> gpurun_out/r2n_bench_$n.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d.get('latency_1stream',{}).get('p50_ms'))
    except Exception as e:
        print(f, 'ERR', e)
PY
