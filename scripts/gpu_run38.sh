cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|parity set|Error|assert" | head
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b38.json 2> gpurun_out/b38.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/b38.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'])"
) > gpurun_out/run38.log 2>&1
cat gpurun_out/run38.log
