cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ni in 2 1; do for w in 14 12 10 8; do TINYMPC_GPS_NI=$ni TINYMPC_GPS_WARPS=$w timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1; done; done > gpurun_out/r02c_c4_ni.txt
timeout 900 python tools/auto_rule_sweep.py --reps 2 > gpurun_out/r02c_auto_sweep.md 2>&1
TINYMPC_GPS_NI=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 1 -c 1 -o gpurun_out/r02c_gps_c4_ni1 python tools/quick_bench.py --config c4 --kernel gps --reps 1 > gpurun_out/r02c_ncu.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "precompute_vs_reference or golden" 2>&1 | tail -5) > gpurun_out/r02c_pytest.txt
cat gpurun_out/r02c_c4_ni.txt; tail -3 gpurun_out/r02c_pytest.txt
