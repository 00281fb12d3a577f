cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for c in 1 0; do
  export PARAKEET_B200_ATTN_CFG=$c
  timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "bf16" 2>&1 | tail -2
  for s in 1024 128; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/b34_$c.json 2> gpurun_out/b34_$c.err; echo "cfg=$c streams=$s rc=$?"; tail -2 gpurun_out/b34_$c.err
  python -c "import json;d=json.load(open('gpurun_out/b34_$c.json'));print(d['ms_per_step'],d['value'])"
  done; done
) > gpurun_out/run34.log 2>&1
cat gpurun_out/run34.log
