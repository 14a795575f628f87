cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or ragged or edge_cases or errors" 2>&1 | tail -15 > gpurun_out/r02_t3.txt
for ni in 1 2; do for w in 8 7; do TINYMPC_GPS_NI=$ni TINYMPC_GPS_WARPS=$w timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1; done; done > gpurun_out/r02_c4_gps3.txt
TINYMPC_GPS_NI=2 TINYMPC_GPS_DIST=1 timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps3.txt
TINYMPC_GPS_NI=2 TINYMPC_GPS_DIST=3 timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps3.txt
timeout 300 python tools/quick_bench.py --config c2 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps3.txt
timeout 300 python tools/quick_bench.py --config c3 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps3.txt
cat gpurun_out/r02_t3.txt gpurun_out/r02_c4_gps3.txt
