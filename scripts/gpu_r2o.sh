# round 2: per-kernel durations of the streaming LayerNorm / dwconv4 (ncu launch list), A/B bench
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_pytest.log
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"layernorm|dwconv" -s 8000 -c 60 --csv --log-file gpurun_out/r2o_ln.csv python bench.py $BA > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2o_ln.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn,mn,mv=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hi+1:]:
    if len(r)>mv: agg[r[kn].split('(')[0]][r[mn]].append(float(r[mv].replace(',','')))
for k,d in agg.items():
    print(k, {m:(round(sum(v)/len(v),2),len(v)) for m,v in d.items()})
PY
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r2o_bench_new.json 2> gpurun_out/r2o_bench_new.err
PARAKEET_B200_LN_STREAM=0 PARAKEET_B200_DWCONV4=0 timeout 600 python bench.py $BB > gpurun_out/r2o_bench_old.json 2> gpurun_out/r2o_bench_old.err
PARAKEET_B200_LN_STREAM=0 timeout 600 python bench.py $BB > gpurun_out/r2o_bench_dw4only.json 2> gpurun_out/r2o_bench_dw4only.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2o_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
