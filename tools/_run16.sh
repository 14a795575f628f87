mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tpi_solve -c 1 -o gpurun_out/r01_tpi_c4_strict python tools/quick_bench.py --kernel tpi --config c4 --mode strict --reps 0 --max_iter 10 > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
