# per-kernel durations for kernels matching $PAT (short prefill; ncu launch list, warm caches)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --prefill-chunks 4 ${EXTRA:-}"
  timeout 300 python bench.py $BA > gpurun_out/k_plain.json 2> gpurun_out/k_plain.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"$PAT" -s ${SKIP:-8} -c ${CNT:-12} --csv --log-file gpurun_out/kern.csv python bench.py $BA > gpurun_out/k_ncu.log 2>&1
  echo "ncu rc=$?"
) > gpurun_out/kern.log 2>&1
tail -2 gpurun_out/kern.log
