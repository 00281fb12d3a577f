cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-latency --steps 5 --prefill-chunks 4 --longform ${LF:-32} > gpurun_out/longform.json 2> gpurun_out/longform.err; echo "rc=$?"
tail -2 gpurun_out/longform.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/longform.json').read().strip().splitlines()[-1])
print(d['longform'])
PY
