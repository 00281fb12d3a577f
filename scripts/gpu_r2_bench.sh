# round-2: targeted tests, then bench A/B (graph on/off) at 1024 and 128 streams
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine_api.py tests/test_gpu_legacy_abi.py -m gpu -q -x -k "graph or running or prologue" > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench_default.err
PARAKEET_B200_GRAPH=0 timeout 600 python bench.py --no-cpu-baseline --no-latency --no-config3 > gpurun_out/r2b_bench_nograph.json 2> gpurun_out/r2b_bench_nograph.err
for n in 128 256; do
  timeout 600 python bench.py --streams $n --no-cpu-baseline --no-latency --no-config3 > gpurun_out/r2b_bench_${n}.json 2> gpurun_out/r2b_bench_${n}.err
  PARAKEET_B200_GRAPH=0 timeout 600 python bench.py --streams $n --no-cpu-baseline --no-latency --no-config3 > gpurun_out/r2b_bench_${n}_nograph.json 2> gpurun_out/r2b_bench_${n}_nograph.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'graphs', d['config'].get('step_graphs'), 'roof', round(d['roofline']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e)
PY
