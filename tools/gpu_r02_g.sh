cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r02g_pytest.txt
for c in c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 4 2>&1 | tail -1; done > gpurun_out/r02g_quick.txt
TINYMPC_GPI_TMEM=0 timeout 300 python tools/quick_bench.py --config c2 --kernel gpi --reps 3 2>&1 | tail -1 >> gpurun_out/r02g_quick.txt
timeout 300 python tools/quick_bench.py --config c2 --kernel gpi --mode fast --reps 3 2>&1 | tail -1 >> gpurun_out/r02g_quick.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 1 -c 1 -o gpurun_out/r02g_gpi_c2 python tools/quick_bench.py --config c2 --kernel gpi --reps 1 > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_pytest.txt; cat gpurun_out/r02g_quick.txt
