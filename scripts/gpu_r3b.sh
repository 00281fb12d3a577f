# round 2: CTA-pair split-K for the long-K residual GEMMs at small batches: mid-size parity test, A/B at 64 ... 320 streams
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_saturated_parity.py -m gpu -q -x -k "mid_size or saturated" > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3b_pytest.log; tail -4 gpurun_out/r3b_pytest.log
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
for n in 64 128 192 256 320; do
  timeout 600 python bench.py --streams $n $BB > gpurun_out/r3b_bench_$n.json 2> gpurun_out/r3b_bench_$n.err
  PARAKEET_B200_PAIR_SMALL=0 timeout 600 python bench.py --streams $n $BB > gpurun_out/r3b_bench_${n}_off.json 2> gpurun_out/r3b_bench_${n}_off.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
