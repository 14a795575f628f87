# round-2 GPU session B: new parity tests, compressed library, GPS after predication / deeper backward ring / fp64 absmax
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25) > gpurun_out/r02b_pytest.txt
for c in c4 c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 3 2>&1 | tail -1; done > gpurun_out/r02b_quick.txt
for k in gps tpi; do timeout 300 python tools/quick_bench.py --config c2 --kernel $k --reps 3 2>&1 | tail -1; done >> gpurun_out/r02b_quick.txt
timeout 300 python tools/quick_bench.py --config c3 --kernel gps --reps 3 2>&1 | tail -1 >> gpurun_out/r02b_quick.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 1 -c 1 -o gpurun_out/r02b_gps_c4 python tools/quick_bench.py --config c4 --kernel gps --reps 1 > gpurun_out/r02b_ncu_gps_c4.log 2>&1
tail -4 gpurun_out/r02b_pytest.txt; cat gpurun_out/r02b_quick.txt
