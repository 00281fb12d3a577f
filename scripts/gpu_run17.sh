cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for s in 128 256 512; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/b17_$s.json 2> gpurun_out/b17_$s.err; echo "streams=$s rc=$?"
    python -c "import json;d=json.load(open('gpurun_out/b17_$s.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'],d['roofline']['achieved'])"
  done
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 64 > gpurun_out/b17_lat.json 2> gpurun_out/b17_lat.err
  python -c "import json;d=json.load(open('gpurun_out/b17_lat.json'));print(d['ms_per_step'],d['latency_1stream'])"
) > gpurun_out/run17.log 2>&1
tail -12 gpurun_out/run17.log
