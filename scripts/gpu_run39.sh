cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x -s -k "bf16 or precision0 or many or offline" 2>&1 | grep -E "passed|failed|parity set|Error|assert" | head
  for m in 1 0; do
  export PARAKEET_B200_SUB_MMA=$m
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b39_$m.json 2> gpurun_out/b39_$m.err; echo "sub_mma=$m rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/b39_$m.json'));print(d['ms_per_step'],d['value'])"
  done
) > gpurun_out/run39.log 2>&1
cat gpurun_out/run39.log
