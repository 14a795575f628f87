set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/gpi_c2_strict python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 --max_iter 20 > gpurun_out/ncu_gpi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tpi_solve -c 1 -o gpurun_out/tpi_c2_fast python tools/quick_bench.py --kernel tpi --config c2 --mode fast --reps 0 --max_iter 20 > gpurun_out/ncu_tpi.log 2>&1
ls -la gpurun_out
