# round 2: packed-math LayerNorm, LN fused into the 1-stream GEMMs: full GPU suite, bench (with latency), A/B, kernel durations
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -5 gpurun_out/r2r_pytest.log
BB="--no-cpu-baseline --no-config3 --longform 0"
timeout 600 python bench.py $BB > gpurun_out/r2r_bench_new.json 2> gpurun_out/r2r_bench_new.err
PARAKEET_B200_LN_FUSE=0 timeout 600 python bench.py $BB > gpurun_out/r2r_bench_nofuse.json 2> gpurun_out/r2r_bench_nofuse.err
for n in 128 256; do timeout 600 python bench.py --streams $n $BB --no-latency > gpurun_out/r2r_bench_$n.json 2> gpurun_out/r2r_bench_$n.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d.get('latency_1stream'))
    except Exception as e:
        print(f, 'ERR', e)
PY
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"layernorm|dwconv" -s 8000 -c 60 --csv --log-file gpurun_out/r2r_ln.csv python bench.py $BA > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2r_ln.csv
PARAKEET_B200_GRAPH=0 CHUNKS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2r_launches_1stream.csv python scripts/probe_1stream.py > gpurun_out/r2r_ncu1s.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2r_launches_1stream.csv > gpurun_out/r2r_launch_summary_1stream.csv 2>&1; head -14 gpurun_out/r2r_launch_summary_1stream.csv
