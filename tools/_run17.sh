mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; cat gpurun_out/bench_r01_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')}); print(d['roofline']); print(d['cpu_baseline'])"; tail -2 gpurun_out/bench_r01_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/r01_gpi_final_c2_strict python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tpi_solve -c 1 -o gpurun_out/r01_tpi_final_c2_strict python tools/quick_bench.py --kernel tpi --config c2 --mode strict --reps 0 > gpurun_out/ncu_f.log 2>&1
