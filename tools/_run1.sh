set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or ragged or edge_cases" 2>&1 | tail -15 > gpurun_out/r02_t1.txt
for d in 1 2 3; do for w in 8 7 4; do TINYMPC_GPS_DIST=$d TINYMPC_GPS_WARPS=$w timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1; done; done > gpurun_out/r02_c4_gps.txt
timeout 300 python tools/quick_bench.py --config c4 --kernel tpi --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps.txt
timeout 300 python tools/quick_bench.py --config c2 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps.txt
timeout 300 python tools/quick_bench.py --config c3 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps.txt
timeout 300 python tools/quick_bench.py --config c2 --kernel gpi --reps 3 2>&1 | tail -1 >> gpurun_out/r02_c4_gps.txt
cat gpurun_out/r02_t1.txt gpurun_out/r02_c4_gps.txt
