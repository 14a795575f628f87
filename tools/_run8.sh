mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; tail -c 600 gpurun_out/bench_r01_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/r01_gpi_c2_strict python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tpi_solve -c 1 -o gpurun_out/r01_tpi_c2_strict python tools/quick_bench.py --kernel tpi --config c2 --mode strict --reps 0 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -8
