# round 2: two half-batch engines per GPU at 1024 streams (HBM-bound attention of one chain under the tensor-bound GEMMs of the other?)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r3c_bench_1eng.json 2> gpurun_out/r3c_bench_1eng.err
timeout 600 python bench.py $BB --engines-per-gpu 2 > gpurun_out/r3c_bench_2eng.json 2> gpurun_out/r3c_bench_2eng.err
PARAKEET_B200_GEMM_MAX_PAIRS=56 timeout 600 python bench.py $BB --engines-per-gpu 2 > gpurun_out/r3c_bench_2eng_56.json 2> gpurun_out/r3c_bench_2eng_56.err
timeout 600 python bench.py $BB --engines-per-gpu 3 > gpurun_out/r3c_bench_3eng.json 2> gpurun_out/r3c_bench_3eng.err
PARAKEET_B200_PAIR_MIN_M=1024 timeout 600 python bench.py --streams 256 $BB > gpurun_out/r3c_bench_256_pairffn.json 2> gpurun_out/r3c_bench_256_pairffn.err
timeout 600 python bench.py --streams 256 $BB > gpurun_out/r3c_bench_256.json 2> gpurun_out/r3c_bench_256.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3c_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
