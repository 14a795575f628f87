#!/bin/bash
# Regenerates the raw material of profiles/ on a B200 box:  gpurun --timeout 1500 -- bash tools/profile_refresh.sh
# (outputs land in gpurun_out/; the summaries under profiles/ are written from them by hand / tools).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
T="timeout 600"
$T python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 3 -c 1 -f -o gpurun_out/r01_gpi_tmem_c2_strict \
    python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 > gpurun_out/ncu_gpi_tmem.log 2>&1
$T python tools/roofline_sweep.py > gpurun_out/sweep.md 2> gpurun_out/sweep.err
$T python tools/closed_loop_bench.py > gpurun_out/closed_loop.txt 2>&1
for k in gpi tpi; do $T python tools/quick_bench.py --kernel $k --config c3 --mode strict --reps 5 2>&1 | tail -1; done > gpurun_out/c3.txt
$T python tools/quick_bench.py --kernel gpi --config c3 --mode fast --reps 5 2>&1 | tail -1 >> gpurun_out/c3.txt
tail -c 600 gpurun_out/bench_n1.json; tail -3 gpurun_out/closed_loop.txt; cat gpurun_out/c3.txt; tail -5 gpurun_out/sweep.md
