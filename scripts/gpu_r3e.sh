# persistent decode with register-blocked dots: whole-utterance parity tests, config-5 shape at 4 and 32 clips
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3e_pytest.log; tail -3 gpurun_out/r3e_pytest.log
for n in 4 32; do
  timeout 1500 python bench.py --no-cpu-baseline --no-config3 --steps 5 --longform $n > gpurun_out/r3e_bench_lf$n.json 2> gpurun_out/r3e_bench_lf$n.err; echo "rc=$?"
done
PARAKEET_B200_DECODE_PERSIST=0 timeout 1500 python bench.py --no-cpu-baseline --no-config3 --steps 5 --longform 32 > gpurun_out/r3e_bench_lf32_graph.json 2> gpurun_out/r3e_bench_lf32_graph.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3e_bench_lf*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    c=d.get('config5_longform') or {}
    print(f, {k:c.get(k) for k in ('wall_s','rtfx_e2e','tokens','decode_ms','attention_ms','gemm_ms')})
PY
