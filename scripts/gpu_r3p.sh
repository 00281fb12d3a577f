# round 2, final evidence set of the tree: launch lists (1024 / 128 / 1 streams), ncu --set full of the hot kernels, sanitizer pass on the new kernels
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
timeout 600 python bench.py $BA > gpurun_out/r3p_plain.json 2> gpurun_out/r3p_plain.err; echo "plain rc=$?"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 36000 -c 1200 --csv --log-file gpurun_out/r02f_launches_1024.csv python bench.py $BA > gpurun_out/r3p_ncu1.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r02f_launches_1024.csv > gpurun_out/r02f_launch_summary_1024.csv 2>&1; head -26 gpurun_out/r02f_launch_summary_1024.csv
PARAKEET_B200_GRAPH=0 STREAMS=128 CHUNKS=92 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 33000 -c 1100 --csv --log-file gpurun_out/r02f_launches_128.csv python scripts/probe_1stream.py > gpurun_out/r3p_ncu128.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r02f_launches_128.csv > gpurun_out/r02f_launch_summary_128.csv 2>&1; head -12 gpurun_out/r02f_launch_summary_128.csv
PARAKEET_B200_GRAPH=0 CHUNKS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02f_launches_1stream.csv python scripts/probe_1stream.py > gpurun_out/r3p_ncu1s.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r02f_launches_1stream.csv > gpurun_out/r02f_launch_summary_1stream.csv 2>&1; head -10 gpurun_out/r02f_launch_summary_1stream.csv
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 19200 -c 14 -o gpurun_out/r02f_gemm -f python bench.py $BA > gpurun_out/r3p_ncu2.log 2>&1; echo "ncu gemm rc=$?"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none -k regex:"layernorm|dwconv|attention_mma" -s 3000 -c 8 -o gpurun_out/r02f_misc -f python bench.py $BA > gpurun_out/r3p_ncu3.log 2>&1; echo "ncu misc rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"decode_persistent" -c 1 -o gpurun_out/r02f_decode_persistent -f python scripts/lf_decode_probe.py 120 2 4 > gpurun_out/r3p_ncu4.log 2>&1; echo "ncu persist rc=$?"
for f in r02f_gemm r02f_misc r02f_decode_persistent; do python scripts/ncu_summary.py full gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_full_summary.txt 2>&1; done
ls -la gpurun_out/r02f*.ncu-rep
