cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r02h_pytest.txt
for c in c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 4 2>&1 | tail -1; done > gpurun_out/r02h_quick.txt
TINYMPC_GPI_TMEM=0 timeout 300 python tools/quick_bench.py --config c2 --kernel gpi --reps 3 2>&1 | tail -1 >> gpurun_out/r02h_quick.txt
for m in 0 1 3 4 7 15 20 23 31 39 55; do echo "L2MODE $m"; TINYMPC_GPS_L2=$m timeout 300 python tools/quick_bench.py --config c4 --kernel gps --reps 3 2>&1 | tail -1 | cut -c1-200; done > gpurun_out/r02h_l2modes.txt 2>&1
tail -3 gpurun_out/r02h_pytest.txt; cat gpurun_out/r02h_quick.txt | cut -c1-200; cat gpurun_out/r02h_l2modes.txt
