# round 2: pair threshold check at 512 / 640 streams, then the whole GPU suite, smoke, default bench line, reference arm
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
for n in 512 640; do timeout 600 python bench.py --streams $n $BB > gpurun_out/r2z_bench_$n.json 2> gpurun_out/r2z_bench_$n.err; done
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -6 gpurun_out/r2z_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log; tail -3 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d.get('roofline',{}).get('frac'), d.get('clocks',{}).get('sm_mhz'), d.get('latency_1stream',{}).get('p50_ms'), (d.get('config5_longform') or {}).get('wall_s'))
    except Exception as e:
        print(f, 'ERR', e)
PY
