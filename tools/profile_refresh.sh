#!/bin/bash
# Regenerates the raw material of profiles/ (round 2) on a B200 box:  gpurun --timeout 3000 -- bash tools/profile_refresh.sh
# Outputs land in gpurun_out/r02_*; the tables of profiles/README.md are printed from them by tools/bench_table.py and
# tools/ncu_summary.py.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
T="timeout 900"
($T python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/r02_pytest.txt
($T python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r02_bench_ref.err | tail -1) > gpurun_out/r02_bench_reference_arm.json
($T python bench.py --steps 10 --warmup 3 2> gpurun_out/r02_bench_n1.err | tail -1) > gpurun_out/r02_bench_n1.json
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 3 -c 1 -f -o gpurun_out/r02_gpi_c2 \
    python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 > gpurun_out/r02_ncu_gpi_c2.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 3 -c 1 -f -o gpurun_out/r02_gps_c4 \
    python tools/quick_bench.py --kernel gps --config c4 --mode strict --reps 3 > gpurun_out/r02_ncu_gps_c4.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 3 -c 1 -f -o gpurun_out/r02_gps_c4p \
    python tools/quick_bench.py --kernel gps --config c4p --mode strict --reps 3 > gpurun_out/r02_ncu_gps_c4p.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 3 -c 1 -f -o gpurun_out/r02_gpi_c3 \
    python tools/quick_bench.py --kernel gpi --config c3 --mode strict --reps 3 > gpurun_out/r02_ncu_gpi_c3.log 2>&1
$T python tools/roofline_sweep.py > gpurun_out/r02_sweep_1gpu.md 2> gpurun_out/r02_sweep.err
$T python tools/auto_rule_sweep.py --reps 1 > gpurun_out/r02_auto_rule_sweep.md 2>&1
$T python tools/closed_loop_bench.py > gpurun_out/r02_closed_loop.txt 2>&1
for k in gpi gps tpi; do for c in c2 c3; do $T python tools/quick_bench.py --kernel $k --config $c --mode strict --reps 4 2>&1 | tail -1; done; done > gpurun_out/r02_quick.txt
for k in gps tpi; do $T python tools/quick_bench.py --kernel $k --config c4 --mode strict --reps 3 2>&1 | tail -1; done >> gpurun_out/r02_quick.txt
$T python tools/quick_bench.py --kernel gpi --config c2 --mode fast --reps 4 2>&1 | tail -1 >> gpurun_out/r02_quick.txt
tail -2 gpurun_out/r02_pytest.txt; cut -c1-300 gpurun_out/r02_bench_n1.json; cut -c1-200 gpurun_out/r02_quick.txt
