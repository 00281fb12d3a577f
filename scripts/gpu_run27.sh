cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for s in 16 64; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --streams $s > gpurun_out/b27_$s.json 2> gpurun_out/b27_$s.err; echo "streams=$s rc=$?"; tail -3 gpurun_out/b27_$s.err
  python -c "import json;d=json.load(open('gpurun_out/b27_$s.json'));print(d['ms_per_step'],d['value'],d['config']['wall_ms_per_step_resident'], d.get('latency_1stream'))"
  done
  BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --streams 128"
  timeout 600 python bench.py $BA > gpurun_out/b27s.json 2> gpurun_out/b27s.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_r1e_128.csv python bench.py $BA > gpurun_out/b27_ncu1.log 2>&1
  echo "ncu1 rc=$?"
) > gpurun_out/run27.log 2>&1
tail -12 gpurun_out/run27.log
