cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for sp in 1 2; do PARAKEET_B200_LF_SPLIT=$sp timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2h_lf_probe_split$sp.log 2>&1; tail -1 gpurun_out/r2h_lf_probe_split$sp.log; done
timeout 900 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -s > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
PARAKEET_B200_LF_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -k "ragged or tiny or tcgen05" > gpurun_out/r2h_pytest_split2.log 2>&1; tail -2 gpurun_out/r2h_pytest_split2.log
timeout 300 python - <<'PY' > gpurun_out/r2h_latency.log 2>&1
import sys, os
sys.path[:0] = ['trt-asr-engine_b200', 'trt-asr-engine_b200/tools']
import numpy as np, bench, binding
from make_synthetic_model import ensure_model
from synth_audio import synth_clip
model = ensure_model('models/synth24', n_layers=24, seed=0)
clip = synth_clip(8.0, 1000)
print(bench.latency_one_stream(binding, model, 0, clip))
PY
tail -2 gpurun_out/r2h_latency.log
PARAKEET_B200_LF_SPLIT=${BEST:-1} timeout 900 ncu --set full --clock-control none --import-source on -k regex:lf_attention_tc -c 1 -o gpurun_out/r02c_lfattn_tc -f python scripts/lf_probe.py 600 2 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py full gpurun_out/r02c_lfattn_tc.ncu-rep > gpurun_out/r02c_lfattn_tc_ncu_full_summary.txt 2>&1
