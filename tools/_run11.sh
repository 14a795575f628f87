mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; tail -c 1200 gpurun_out/bench_r01_n1.json; cat gpurun_out/bench_r01_n1.err | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2>&1; tail -c 600 gpurun_out/bench_r01_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python tools/roofline_sweep.py --reps 2 > gpurun_out/sweep_r01.md 2> gpurun_out/sweep_r01.err; tail -45 gpurun_out/sweep_r01.md; tail -3 gpurun_out/sweep_r01.err
