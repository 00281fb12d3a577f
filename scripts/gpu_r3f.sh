# round 2, final validation of the tree: whole GPU suite, smoke, default bench line, reference arm, BASELINE config 5 exactly
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log; tail -6 gpurun_out/r3f_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3f_smoke.log; tail -3 gpurun_out/r3f_smoke.log
timeout 900 python bench.py > gpurun_out/r3f_bench_default.json 2> gpurun_out/r3f_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_bench_reference.err; echo "ref rc=$?"
timeout 1500 python bench.py --no-cpu-baseline --no-config3 --steps 5 --longform 32 > gpurun_out/r3f_bench_lf32.json 2> gpurun_out/r3f_bench_lf32.err; echo "lf32 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('config5_longform') or {}
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d.get('roofline',{}).get('frac'), d.get('clocks',{}).get('sm_mhz'), d.get('latency_1stream',{}).get('p50_ms'), {k:c.get(k) for k in ('wall_s','decode_ms','attention_ms')}, (d.get('config3_64streams_mixed_cache') or {}).get('ms_per_step_e2e'))
    except Exception as e:
        print(f, 'ERR', e)
PY
