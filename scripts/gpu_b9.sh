cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for CFG in "128 1" "128 2" "256 2" "512 2" "128 4" "256 4"; do
set -- $CFG
timeout 600 python bench.py --no-cpu-baseline --no-latency --streams $1 --engines-per-gpu $2 > gpurun_out/b9_$1_$2.json 2> gpurun_out/b9_$1_$2.err; echo "streams=$1 engines=$2 rc=$?"
tail -1 gpurun_out/b9_$1_$2.err
python - <<PY
import json
d = json.loads(open('gpurun_out/b9_$1_$2.json').read().strip().splitlines()[-1])
print($1, $2, round(d['value']), round(d['ms_per_step'],3), 'wall', round(d['config']['wall_ms_per_step_resident'],3), 'e2e', round(d['e2e']['value']), round(d['roofline']['achieved']))
PY
done
