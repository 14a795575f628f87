mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/r01_gpi_v5_c2_strict python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 --max_iter 25 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
