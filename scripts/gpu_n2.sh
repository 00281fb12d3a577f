cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_n2.json 2> gpurun_out/b_n2.err; echo rc=$?
tail -2 gpurun_out/b_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/b_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['frac'], [round(r['achieved']) for r in d['roofline_hbm']])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
