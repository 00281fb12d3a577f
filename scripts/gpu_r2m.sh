# round 2 (re-entry): validate HEAD end to end, then the r02 evidence set: default bench line, reference arm, launch lists, ncu --set full captures
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2m_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -22 gpurun_out/r2m_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2m_smoke.log; tail -3 gpurun_out/r2m_smoke.log
timeout 900 python bench.py > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r2m_bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err; echo "ref rc=$?"
for n in 128 256; do timeout 600 python bench.py --streams $n --no-cpu-baseline --no-latency --no-config3 --longform 0 > gpurun_out/r2m_bench_$n.json 2> gpurun_out/r2m_bench_$n.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d.get('roofline',{}).get('frac'), d.get('clocks',{}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 36000 -c 1200 --csv --log-file gpurun_out/r02_launches_1024.csv python bench.py $BA > gpurun_out/r2m_ncu1.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r02_launches_1024.csv > gpurun_out/r02_launch_summary_1024.csv 2>&1; head -24 gpurun_out/r02_launch_summary_1024.csv
PARAKEET_B200_GRAPH=0 CHUNKS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_launches_1stream.csv python scripts/probe_1stream.py > gpurun_out/r2m_ncu1s.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r02_launches_1stream.csv > gpurun_out/r02_launch_summary_1stream.csv 2>&1; head -8 gpurun_out/r02_launch_summary_1stream.csv
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 19200 -c 14 -o gpurun_out/r02_gemm -f python bench.py $BA > gpurun_out/r2m_ncu2.log 2>&1; echo "ncu gemm rc=$?"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_mma" -s 2210 -c 1 -o gpurun_out/r02_attn -f python bench.py $BA > gpurun_out/r2m_ncu3.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lf_attention_tc -c 1 -o gpurun_out/r02_lfattn_tc -f python scripts/lf_probe.py 600 2 > gpurun_out/r2m_ncu4.log 2>&1; echo "ncu lf rc=$?"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none -k regex:"logmel|tdt_select|lstm_cell|joint_hidden|pred_input|layernorm|dwconv" -s 400 -c 14 -o gpurun_out/r02_misc -f python bench.py $BA > gpurun_out/r2m_ncu5.log 2>&1; echo "ncu misc rc=$?"
for f in r02_gemm r02_attn r02_lfattn_tc r02_misc; do python scripts/ncu_summary.py full gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_full_summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
