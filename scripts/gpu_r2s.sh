# round 2: probes -- cluster occupancy, L2 -> SM contention of the CTA-pair GEMM (fewer resident pairs), launch list at 128 streams
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
$K clusters > gpurun_out/r2s_clusters.txt 2>&1; cat gpurun_out/r2s_clusters.txt
for mp in 74 56 37 18; do
  echo "== max pairs $mp"; PARAKEET_B200_GEMM_MAX_PAIRS=$mp $K gemm 6144 4096 1024 50 silu 0 8 2>&1 | tail -1
  PARAKEET_B200_GEMM_MAX_PAIRS=$mp $K gemm 6144 1024 4096 50 partial2pb 0 8 2>&1 | tail -1
done > gpurun_out/r2s_maxpairs.txt 2>&1; cat gpurun_out/r2s_maxpairs.txt
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r2s_bench_new.json 2> gpurun_out/r2s_bench_new.err
PARAKEET_B200_PAIR_MODES=80 timeout 600 python bench.py $BB > gpurun_out/r2s_bench_pm80.json 2> gpurun_out/r2s_bench_pm80.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
PARAKEET_B200_GRAPH=0 STREAMS=128 CHUNKS=92 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 38000 -c 1200 --csv --log-file gpurun_out/r2s_launches_128.csv python scripts/probe_1stream.py > gpurun_out/r2s_ncu128.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2s_launches_128.csv > gpurun_out/r02_launch_summary_128.csv 2>&1; head -30 gpurun_out/r02_launch_summary_128.csv
