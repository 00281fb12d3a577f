cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py tests/test_cli.py -m gpu -q -x -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -5 gpurun_out/r2f_pytest.log
PARAKEET_B200_LF_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -k "ragged or tiny or tcgen05" > gpurun_out/r2f_pytest_split2.log 2>&1; tail -2 gpurun_out/r2f_pytest_split2.log
for sp in 1 2; do PARAKEET_B200_LF_SPLIT=$sp timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2f_lf_probe_split$sp.log 2>&1; tail -1 gpurun_out/r2f_lf_probe_split$sp.log; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lf_attention_tc -c 1 -o gpurun_out/r02b_lfattn_tc -f python scripts/lf_probe.py 600 2 > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu lf rc=$?"
python scripts/ncu_summary.py full gpurun_out/r02b_lfattn_tc.ncu-rep > gpurun_out/r02b_lfattn_tc_ncu_full_summary.txt 2>&1
