cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( export PARAKEET_B200_PART_BF16=1
  timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x -s -k "bf16" 2>&1 | grep -E "passed|failed|parity set|Error|assert" | head
  for pb in 1 0; do
  export PARAKEET_B200_PART_BF16=$pb
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b37_$pb.json 2> gpurun_out/b37_$pb.err; echo "part_bf16=$pb rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/b37_$pb.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'])"
  done
) > gpurun_out/run37.log 2>&1
cat gpurun_out/run37.log
