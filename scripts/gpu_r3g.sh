# round 2: mid-size batches (N = 4 point of the scaling curve): pair kernel for FFN-up from fewer rows, streaming LayerNorm from fewer rows
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
for n in 192 256 320; do
  timeout 600 python bench.py --streams $n $BB > gpurun_out/r3g_bench_${n}_a_default.json 2> /dev/null
  PARAKEET_B200_PAIR_MIN_M=1024 timeout 600 python bench.py --streams $n $BB > gpurun_out/r3g_bench_${n}_b_pair1024.json 2> /dev/null
  PARAKEET_B200_LN_STREAM=1024 timeout 600 python bench.py --streams $n $BB > gpurun_out/r3g_bench_${n}_c_ln1024.json 2> /dev/null
  PARAKEET_B200_PAIR_MIN_M=1024 PARAKEET_B200_LN_STREAM=1024 timeout 600 python bench.py --streams $n $BB > gpurun_out/r3g_bench_${n}_d_both.json 2> /dev/null
  timeout 600 python bench.py --streams $n $BB > gpurun_out/r3g_bench_${n}_e_default2.json 2> /dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3g_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']))
    except Exception as e:
        print(f, 'ERR', e)
PY
