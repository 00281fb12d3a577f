# round 2: where does the CTA-pair kernel for every projection pay off? (stream counts between 256 and 1024)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
for n in 320 384 512 640 768 1024; do
  timeout 600 python bench.py --streams $n $BB > gpurun_out/r2y_bench_$n.json 2> gpurun_out/r2y_bench_$n.err
  PARAKEET_B200_GEMM_BN=512 timeout 600 python bench.py --streams $n $BB > gpurun_out/r2y_bench_${n}_pair.json 2> gpurun_out/r2y_bench_${n}_pair.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
