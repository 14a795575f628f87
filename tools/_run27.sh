mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; tail -c 400 gpurun_out/bench_r01_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/r01_gpi_final_c2_strict python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 > gpurun_out/ncu_e.log 2>&1
python tools/roofline_sweep.py --reps 2 > gpurun_out/sweep_r01.md 2> gpurun_out/sweep_r01.err; head -12 gpurun_out/sweep_r01.md | cut -c1-220
