cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gps or precompute_vs_reference or auto or rocket" 2>&1 | tail -8) > gpurun_out/r02e_pytest.txt
for c in c4 c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel gps --reps 3 2>&1 | tail -1; done > gpurun_out/r02e_quick.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 1 -c 1 -o gpurun_out/r02e_gps_c4 python tools/quick_bench.py --config c4 --kernel gps --reps 1 > gpurun_out/r02e_ncu.log 2>&1
timeout 900 python tools/auto_rule_sweep.py --reps 1 > gpurun_out/r02e_auto_sweep.md 2>&1
tail -4 gpurun_out/r02e_pytest.txt; cat gpurun_out/r02e_quick.txt
