cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4) > gpurun_out/r02m3_pytest.txt
for c in c2 c3 c4; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 4 2>&1 | tail -1; done > gpurun_out/r02m3_quick.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 1 -c 1 -o gpurun_out/r02_gpi_f64_12_4_50 python tools/auto_rule_sweep.py --reps 1 --only f64_12_4 > gpurun_out/r02m3_ncu.log 2>&1
tail -2 gpurun_out/r02m3_pytest.txt; cut -c1-200 gpurun_out/r02m3_quick.txt
