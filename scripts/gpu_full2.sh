cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r1k.log
tail -5 gpurun_out/tests_r1k.log
timeout 300 python scripts/lf_probe.py 1200 2 > gpurun_out/lf_probe.log 2>&1; tail -3 gpurun_out/lf_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lf_attention_mma -c 1 -o gpurun_out/prof_lfattn_r1k -f python scripts/lf_probe.py 1200 2 > gpurun_out/lf_ncu.log 2>&1
echo "ncu rc=$?"
