cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for S in 512 256 128 64; do
timeout 600 python bench.py --no-cpu-baseline --no-latency --streams $S > gpurun_out/b8_$S.json 2> gpurun_out/b8_$S.err; echo "streams=$S rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/b8_$S.json').read().strip().splitlines()[-1])
print($S, round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), round(d['roofline_hbm'][0]['achieved']), d['gpu_launches'])
PY
done
