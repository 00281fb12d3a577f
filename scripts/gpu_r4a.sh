# round 2, session 3: valid-slot K/V loads in the streaming attention + register-resident FFT frontend -- parity, A/B, ncu traffic
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ab_switches.py tests/test_gpu_frontend.py tests/test_gpu_saturated_parity.py tests/test_gpu_model.py -m gpu -x -q --durations=8 > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r4a_pytest.log
BA="--no-cpu-baseline --no-latency --no-config3 --longform 0"
for E in 0 1 0 1; do
PARAKEET_B200_ATTN_TRIM=$E timeout 600 python bench.py $BA > gpurun_out/r4a_ab_$E.json 2> gpurun_out/r4a_ab.err; rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/r4a_ab_$E.json').read().strip().splitlines()[-1])
print("TRIM=$E rc=$rc", round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['roofline_hbm'][0]['achieved']), 'logmel', round(d['roofline_hbm'][1]['achieved']), d['clocks']['sm_mhz'])
PY
done
PARAKEET_B200_LOGMEL=0 timeout 300 python scripts/logmel_probe.py 2>&1 | tail -1
timeout 300 python scripts/logmel_probe.py 2>&1 | tail -1
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_mma" -s 2208 -c 2 -o gpurun_out/r02g_attn -f python bench.py --steps 2 --warmup 3 $BA > gpurun_out/r4a_ncu1.log 2>&1; echo "ncu attn rc=$?"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel" -s 1 -c 1 -o gpurun_out/r02g_logmel -f python scripts/logmel_probe.py > gpurun_out/r4a_ncu2.log 2>&1; echo "ncu logmel rc=$?"
for f in r02g_attn r02g_logmel; do python scripts/ncu_summary.py full gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_full_summary.txt 2>&1; done
grep -E "dram__bytes|gpu__time_duration" gpurun_out/r02g_attn_ncu_full_summary.txt | head; grep -E "dram__bytes|gpu__time_duration|bank_conflict" gpurun_out/r02g_logmel_ncu_full_summary.txt | head
ls -la gpurun_out/r02g*.ncu-rep
