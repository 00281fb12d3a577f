# round 2, session 3, final validation of the tree: whole GPU suite, smoke, default bench line, reference arm, launch list of the 1024-stream step
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
T=${TAG:-r4f}
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 36000 -c 1200 --csv --log-file gpurun_out/r02g_launches_1024.csv python bench.py $BA > gpurun_out/${T}_ncu1.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r02g_launches_1024.csv > gpurun_out/r02g_launch_summary_1024.csv 2>&1; head -16 gpurun_out/r02g_launch_summary_1024.csv
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('config5_longform') or {}
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d.get('roofline',{}).get('frac'), d.get('clocks',{}).get('sm_mhz'), d.get('latency_1stream',{}).get('p50_ms'), {k:c.get(k) for k in ('wall_s','decode_ms','attention_ms')}, (d.get('config3_64streams_mixed_cache') or {}).get('ms_per_step_e2e'), [round(x['achieved']) for x in d.get('roofline_hbm',[])])
    except Exception as e:
        print(f, 'ERR', e)
PY
if [ "${LF32:-0}" = "1" ]; then
timeout 1500 python bench.py --no-cpu-baseline --no-config3 --steps 5 --longform 32 > gpurun_out/${T}_bench_lf32.json 2> gpurun_out/${T}_bench_lf32.err; echo "lf32 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_lf32.json').read().strip().splitlines()[-1])
print('lf32', {k:v for k,v in d['config5_longform'].items() if k not in ('blank_penalty_calibration','workload')})
PY
fi
