# BASELINE config 5 exactly: 32 clips x 1 h through the whole-utterance path (groups of 4, one batched decode)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python bench.py --no-cpu-baseline --no-config3 --steps 5 --longform 32 > gpurun_out/r3d_bench_lf32.json 2> gpurun_out/r3d_bench_lf32.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d_bench_lf32.json').read().strip().splitlines()[-1])
print(d.get('config5_longform'))
PY
tail -3 gpurun_out/r3d_bench_lf32.err
